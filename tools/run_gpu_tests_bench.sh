#!/bin/bash
# usage: bash tools/run_gpu_tests_bench.sh <tag>  -- GPU tests + the default bench line (kernels unchanged: no ncu pass)
TAG=${1:-x}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -3 ${O}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err
python - ${O}_bench.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value']), "launches", d['gpu_launches'])
for k in ("single_stream","config0","full_ensemble","snr_cfo_batch"):
    print(k, round(d[k]['frames_per_s']), round(d[k]['ms_per_step'],3))
print("fe ok", d['full_ensemble']['payload_equals_transmitted'], "c0 ok", d['config0']['payload_equals_transmitted'])
PY
tail -3 ${O}_bench.err
