#!/bin/bash
# builds build/exp/libzigma_exp$1.so with -DZG_SCAN_EXP=$1 (timing experiments of the scan kernel; results are wrong by design)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/exp/obj$1
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DZG_SCAN_EXP=$1"
for f in zigma_b200/csrc/*.cu; do
  o=build/exp/obj$1/$(basename ${f%.cu}).o
  case $(basename $f) in scan_fwd_bf16.cu) nvcc $FLAGS -c $f -o $o & ;; *) cp build/obj/$(basename ${f%.cu}).o $o ;; esac
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o zigma_b200/lib/libzigma_exp$1.so build/exp/obj$1/*.o -lcudart -lcuda
echo built zigma_b200/lib/libzigma_exp$1.so
