#!/bin/bash
# usage (on the GPU box): bash tools/run_ablate.sh  -- headline bench with parts of the demapper switched off (results are WRONG, times only)
cd $GRAFT_REPO_ROOT
cp dabstar_b200/libdabstar_b200.so /tmp/keep.so
cp dabstar_b200/libdabstar_b200_abl.so dabstar_b200/libdabstar_b200.so
for a in 0 1 2 4 8 3 9 11 15; do
  DABSTAR_DEMAP_ABLATE_AFTER=2 DABSTAR_DEMAP_ABLATE=$a timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-viterbi-sweep --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages']
print('ablate $a: step', round(d['ms_per_step'],2), 'fft', round(s['ingest_fft']['ms_per_step'],2), 'demap', round(s['demap']['ms_per_step'],2))"
done | tee gpurun_out/ablate.log
cp /tmp/keep.so dabstar_b200/libdabstar_b200.so
