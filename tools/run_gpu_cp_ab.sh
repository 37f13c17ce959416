#!/bin/bash
# usage: bash tools/run_gpu_cp_ab.sh <tag>  -- GPU parity tests, then the headline bench with the direct-load and the bulk-copy CP correlation
TAG=${1:-cp}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -4 ${O}_tests.log
for b in 0 1; do
  DABSTAR_CP_BULK=$b timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-viterbi-sweep > ${O}_bench_cp$b.json 2> ${O}_bench_cp$b.err
  python - ${O}_bench_cp$b.json $b <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("cp bulk", sys.argv[2], "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), " ".join(f"{k}={v['ms_per_step']:.3f}" for k, v in d["stages"].items()), "crc", d["run"]["fib_crc_pass"])
PY
  tail -2 ${O}_bench_cp$b.err
done
