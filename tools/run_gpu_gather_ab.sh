#!/bin/bash
# usage: bash tools/run_gpu_gather_ab.sh <tag>  -- GPU parity tests, then the full-ensemble decode with 1 and 2 code words per warp in flight in k_vit_gather
TAG=${1:-ab}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 600 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -4 ${O}_tests.log
for b in 1 2; do
  DABSTAR_GATHER_BATCH=$b timeout 300 python tools/bench_full_ensemble.py > ${O}_fe_batch$b.json 2> ${O}_fe_batch$b.err
  python -c "import json,sys; d=json.load(open('${O}_fe_batch$b.json')); print('batch $b', round(d['ms_per_step'],2), 'ms', round(d['frames_per_s']), 'frames/s', 'msc', round(d['stages_ms']['msc_viterbi'],2), 'fic', round(d['stages_ms']['fic_viterbi'],3), 'ok', d['payload_equals_transmitted'])"
  tail -2 ${O}_fe_batch$b.err
done
