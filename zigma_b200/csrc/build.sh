#!/bin/bash
# Builds libzigma_b200.so (sm_100a) in-tree: zigma_b200/lib/libzigma_b200.so
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
OBJ="$HERE/../../build/obj"
mkdir -p "$OUT" "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
pids=()
for f in "$HERE"/*.cu; do
    o="$OBJ/$(basename "${f%.cu}").o"
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$HERE" -name '*.cuh' -newer "$o")" ] || [ "$HERE/../../include/zigma_b200.h" -nt "$o" ]; then
        ( $NVCC $FLAGS -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; exit 1; } ) &
        pids+=($!)
    fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libzigma_b200.so" "$OBJ"/*.o -lcudart -lcuda
echo "built $OUT/libzigma_b200.so"
