#!/bin/bash
# usage: bash tools/run_gpu_profile.sh <tag>  -- the evidence set of a round: tests, bench (with cpu_baseline), reference arm,
# ncu launch list of the bench command, ncu full capture of the hot kernels, smoke
TAG=${1:-x}
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > ${O}_smi.csv
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; echo "tests exit $?" >> ${O}_tests.log
tail -4 ${O}_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; tail -c 400 ${O}_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2> ${O}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file ${O}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-viterbi-sweep --no-extras > ${O}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_frames|k_demap5|k_vit_tpc|k_vit_gather|k_fic_post|k_cp_corr|k_dip_search|k_prs_corr' -c 14 -o ${O}_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-viterbi-sweep --no-extras > ${O}_ncu_full.log 2>&1
tail -2 ${O}_ncu_full.log
python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log
