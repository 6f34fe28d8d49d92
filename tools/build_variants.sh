#!/bin/bash
# Builds kernel-experiment variants of the C-ABI library into build/variants/<name>.so (select one at run time with
# KBBQ_B200_LIB=...).  usage: tools/build_variants.sh name "EXTRA NVCC FLAGS" [srcdir]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; extra=$2; src=${3:-$ROOT/kbbq_b200/csrc}
out=$ROOT/build/variants; mkdir -p $out/obj_$name
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC \
   --expt-relaxed-constexpr $extra -c -o $out/obj_$name/capi.o $src/kbbq_capi.cu
longobj=
if [ -f $src/kbbq_long.cu ]; then
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC \
     --expt-relaxed-constexpr $extra -c -o $out/obj_$name/long.o $src/kbbq_long.cu
  longobj=$out/obj_$name/long.o
fi
/usr/bin/g++ -std=c++17 -O2 -fPIC -ffp-contract=off -pthread -c -o $out/obj_$name/host.o $src/kbbq_host.cc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -shared -o $out/$name.so $out/obj_$name/capi.o $longobj $out/obj_$name/host.o -Xcompiler -pthread
rm -rf $out/obj_$name
echo built $out/$name.so
