#!/bin/bash
# usage: bash tools/run_gpu_bench.sh <tag> [bench args...]  -- bench only (stdout -> gpurun_out/<tag>_bench.json)
TAG=${1:-x}; shift
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
( time timeout 900 python bench.py "$@" ) > ${O}_bench.json 2> ${O}_bench.err
python - ${O}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d['value']), "ms/step", round(d['ms_per_step'],3), "e2e", round(d['e2e']['value']), "launches", d['gpu_launches'])
print(" ".join(f"{k}={v['ms_per_step']:.3f}" for k,v in d['stages'].items()))
for k in ("two_batches_in_flight","h2d_control","single_stream","snr_cfo_batch","config0","full_ensemble"):
    if k in d: print(k, json.dumps(d[k])[:900])
if d.get("viterbi_sweep"): print("sweep", d["viterbi_sweep"]["mbit_s_overall"], d["viterbi_sweep"]["bits_equal_reference"], {k: round(v["frac_int_alu"],3) for k,v in d["viterbi_sweep"]["levels"].items()})
print("cpu", d.get("cpu_baseline"))
PY
tail -8 ${O}_bench.err
