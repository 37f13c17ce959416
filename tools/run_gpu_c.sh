set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/l_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/l_tests.log
tail -15 gpurun_out/l_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; tail -c 1200 gpurun_out/l_bench.json; tail -5 gpurun_out/l_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_vit_tpc|k_vit_gather|k_fic_post' -c 6 --csv --log-file gpurun_out/l_vit.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -v "^==" gpurun_out/l_vit.csv | cut -d, -f5,13- | head -20
