#!/bin/bash
# usage: [AB_VAR=DABSTAR_TPC_PIPE AB_VALUES="0 1"] bash tools/run_gpu_gather_ab.sh <tag>  -- GPU parity tests, then the bench once per value of a switch
TAG=${1:-ab}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
timeout 900 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -4 ${O}_tests.log
for b in ${AB_VALUES:-2 0}; do
  env ${AB_VAR:-DABSTAR_GATHER_BATCH}=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > ${O}_bench_gather$b.json 2> ${O}_bench_gather$b.err
  python - ${O}_bench_gather$b.json $b <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
fe, c0, sw = d["full_ensemble"], d["config0"], d["viterbi_sweep"]
print("gather", sys.argv[2], "value", round(d["value"]), "fic", round(d["stages"]["fic_viterbi"]["ms_per_step"], 3),
      "| full ensemble", round(fe["frames_per_s"]), "msc", round(fe["stages_ms"]["msc_viterbi"], 2), "gather", round(fe["msc_gather_ms"], 2), "trellis", round(fe["msc_trellis_ms"], 2), fe["payload_equals_transmitted"],
      "| config0", round(c0["frames_per_s"]), "| sweep", round(sw["mbit_s_overall"]), sw["bits_equal_reference"], {k: round(v["frac_int_alu"], 3) for k, v in list(sw["levels"].items())[::4]})
PY
  tail -2 ${O}_bench_gather$b.err
done
