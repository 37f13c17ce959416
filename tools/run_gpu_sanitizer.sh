#!/bin/bash
# usage: bash tools/run_gpu_sanitizer.sh <tag>  -- compute-sanitizer memcheck over the GPU tests of the stages, the chain and the golden vectors
# (the long-recording cases are left out: minutes under the tool), racecheck over the Viterbi / CP-correlation cases
TAG=${1:-san}
cd $GRAFT_REPO_ROOT
O=gpurun_out/${TAG}
{
  echo "# memcheck"
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stages.py tests/test_gpu_chain.py tests/test_gpu_golden.py tests/test_dabplus.py tests/test_tii.py -m gpu -x -q 2>&1 | tail -6
  echo "exit $?"
  echo "# racecheck"
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_golden.py -m gpu -x -q -k "viterbi or cp_correlation or sweep or msc" 2>&1 | tail -12
  echo "exit $?"
} > ${O}_sanitizer.txt 2>&1
cat ${O}_sanitizer.txt | tail -30
