#!/bin/bash
# usage: bash tools/run_gpu.sh <tag> [ncu kernel regex] [extra bench args]   -- tests, bench, optional ncu full capture
TAG=${1:-x}; KREGEX=${2:-}; shift; shift || true
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -6 ${O}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > ${O}_bench.json 2> ${O}_bench.err
python - ${O}_bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("value", d['value'], "ms/step", d['ms_per_step'], "e2e", d['e2e']['value'], "launches", d['gpu_launches'])
for k,v in d['stages'].items(): print(" ", k, round(v['ms_per_step'],4))
PY
tail -5 ${O}_bench.err
if [ -n "$KREGEX" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -c 4 -o ${O}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > ${O}_ncu_full.log 2>&1
  tail -3 ${O}_ncu_full.log
fi
