#!/bin/bash
# Build a variant of libqmcb.so with extra nvcc flags into isingmontecarlo_b200/_variants/<name>/libqmcb.so (select it with
# QMCB_LIB=...).  Usage: tools/build_variant.sh <name> <extra nvcc flags...>
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/isingmontecarlo_b200/_variants/$name
mkdir -p "$out/obj"
cd "$root/isingmontecarlo_b200/csrc"
for f in api sse_serial sse_fast sse_counter sse_rvb classical classical_ref pt; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ --fmad=false "$@" -c $f.cu -o "$out/obj/$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libqmcb.so" "$out"/obj/*.o -ldl
rm -rf "$out/obj"
echo "$out/libqmcb.so"
