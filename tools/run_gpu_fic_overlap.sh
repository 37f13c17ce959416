#!/bin/bash
# usage: bash tools/run_gpu_fic_overlap.sh <tag>  -- GPU parity tests with every window chunked (FIC decoder of a chunk beside the demapper of the next),
# then the headline bench with 1 / 2 / 4 chunks
TAG=${1:-fo}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
DABSTAR_CHUNKS=2 DABSTAR_CHUNK_MODE=1 DABSTAR_CHUNK_MIN_FRAMES=4 timeout 900 python -m pytest tests -m gpu -x -q > ${O}_tests_chunked.log 2>&1; echo "tests exit $?" >> ${O}_tests_chunked.log
tail -3 ${O}_tests_chunked.log
for n in 1 2 4 ${EXTRA_CHUNKS}; do
  DABSTAR_CHUNKS=$n DABSTAR_CHUNK_MODE=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-viterbi-sweep --no-extras > ${O}_bench_chunks$n.json 2> ${O}_bench_chunks$n.err
  python - ${O}_bench_chunks$n.json $n <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("chunks", sys.argv[2], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "crc", d["run"]["fib_crc_pass"], " ".join(f"{k}={v['ms_per_step']:.3f}" for k,v in d['stages'].items() if 'ms_per_step' in v))
PY
done
