#!/bin/bash
# Measurement build of the CUDA library with the demapper's ablation switches (-DDABSTAR_ABLATE): written to gpurun_out-independent
# path dabstar_b200/libdabstar_b200_abl.so; tools/run_ablate.sh swaps it in on the GPU box for the duration of one bench run.
set -e
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DDABSTAR_NO_FAST_MATH -DDABSTAR_ABLATE -Xcompiler -fPIC -cudart shared -I include -I dabstar_b200/csrc"
mkdir -p /tmp/abl_obj
for f in dabstar_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  if [ "$b" = "ofdm_kernels" ] || [ ! -f /tmp/abl_obj/$b.o ] || [ $f -nt /tmp/abl_obj/$b.o ]; then $NV -c -o /tmp/abl_obj/$b.o $f & fi
done
wait
nvcc -shared -cudart shared -o dabstar_b200/libdabstar_b200_abl.so /tmp/abl_obj/*.o
echo built dabstar_b200/libdabstar_b200_abl.so
