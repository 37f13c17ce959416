#!/bin/bash
# usage: bash tools/run_gpu_gather_ncu.sh <tag>  -- ncu --set full of the gather launches of one full-ensemble step (16 recordings keep the replay short)
TAG=${1:-g}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vit_gather --launch-skip 6 -c 6 -o ${O}_gather -f \
  python tools/bench_full_ensemble.py --recordings 96 --frames 104 --steps 1 > ${O}_ncu.log 2>&1
tail -3 ${O}_ncu.log
ls -la ${O}_gather.ncu-rep
