#!/bin/bash
# usage: bash tools/run_gpu_spin_test.sh  -- host control per step against the pool's polling time, the process confined to four CPUs
cd $GRAFT_REPO_ROOT
for spin in 1000 300 100 0; do
  DABSTAR_HOST_SPIN_US=$spin taskset -c 0-3 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --no-viterbi-sweep > gpurun_out/r2_bb_spin$spin.json 2> gpurun_out/r2_bb_spin$spin.err
  python - gpurun_out/r2_bb_spin$spin.json $spin <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("4 cpus, spin_us", sys.argv[2], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "host", round(d["run"]["host_control_ms_per_step"],3))
PY
done
