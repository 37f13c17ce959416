#!/bin/bash
# usage: bash tools/run_gpu_tpc_ncu.sh <tag>  -- ncu --set full of the gather and trellis launches (k_vit_gather*, k_vit_tpc) of the second full-ensemble MSC pass
TAG=${1:-t}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_vit_tpc|k_vit_gather" --launch-skip 14 -c 14 -o ${O}_tpc -f \
  python tools/bench_full_ensemble.py --recordings 96 --frames 104 --steps 1 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log | cut -c1-300
ls -la ${O}_tpc.ncu-rep
