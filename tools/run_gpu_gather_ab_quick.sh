#!/bin/bash
# usage: [AB_VALUES="-1 0"] bash tools/run_gpu_gather_ab_quick.sh <tag>  -- Viterbi / MSC parity tests, then tools/bench_full_ensemble.py once per value of DABSTAR_GATHER_BATCH
TAG=${1:-ab}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_golden.py -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -3 ${O}_tests.log
for b in ${AB_VALUES:--1 0}; do
  DABSTAR_GATHER_BATCH=$b timeout 300 python tools/bench_full_ensemble.py --steps 3 > ${O}_fe_gather$b.json 2> ${O}_fe_gather$b.err
  python - ${O}_fe_gather$b.json $b <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("gather", sys.argv[2], round(d["frames_per_s"]), "ms", round(d["ms_per_step"],2), "msc", round(d["stages_ms"]["msc_viterbi"],2), {k: round(v,2) for k,v in d.items() if k.startswith("msc_")}, d["payload_equals_transmitted"])
PY
done
