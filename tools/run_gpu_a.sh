set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/j_tests.log
tail -5 gpurun_out/j_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; tail -c 3000 gpurun_out/j_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j_bench_ref.json 2> gpurun_out/j_bench_ref.err; tail -c 1500 gpurun_out/j_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_fft_frames|k_demap3|k_viterbi|k_cp_corr' -c 16 -o gpurun_out/j_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j_ncu_full.log 2>&1
ls -la gpurun_out
cat MEASURED_PEAKS.json
