set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/k_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/k_tests.log
tail -15 gpurun_out/k_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; tail -c 2500 gpurun_out/k_bench.json; tail -5 gpurun_out/k_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_vit_tpc|k_vit_gather|k_fic_post' -c 6 -o gpurun_out/k_prof python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/k_ncu_full.log 2>&1
tail -3 gpurun_out/k_ncu_full.log
