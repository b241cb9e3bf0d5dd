#!/bin/bash
# Builds libvit2spn.so (sm_100a only) in-tree.  Usage: ./build.sh [-j N]
set -e
cd "$(dirname "$0")"
SRC="vit-2spn_b200/csrc"
OUT="vit-2spn_b200/libvit2spn.so"
OBJ="build/obj"
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v --expt-relaxed-constexpr"
pids=()
for f in $SRC/*.cu; do
  o="$OBJ/$(basename ${f%.cu}).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC include -name '*.cuh' -newer "$o" -o -name '*.h' -newer "$o" | head -1)" ]; then
    ( $NVCC $FLAGS -c "$f" -o "$o" > "$o.log" 2>&1 || { cat "$o.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" $OBJ/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
