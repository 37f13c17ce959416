#!/bin/bash
# usage: bash tools/run_gpu_quick.sh <tag> [env assignments...]  -- demap/chain parity tests + bench (no cpu baseline)
TAG=${1:-x}; shift
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
for e in "$@"; do export "$e"; done
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -3 ${O}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err
python - ${O}_bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value']), "launches", d['gpu_launches'])
print(" ".join(f"{k}={v['ms_per_step']:.3f}" for k,v in d['stages'].items()))
PY
tail -3 ${O}_bench.err
