#!/bin/bash
# development helper: tools/build_variant.sh NAME "-DFLAG=..."  ->  variants/lib_NAME.so (the product library with extra compile flags)
# on the GPU box: cp variants/lib_NAME.so compu_b200/libcompu_b200.so before a bench run to compare kernel variants in ONE gpurun call
set -e
cd "$(dirname "$0")/../compu_b200/csrc"
name=$1; shift
mkdir -p ../../variants/obj_$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in inflate host deflate synth; do
  /usr/local/cuda/bin/nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp,-Wall,-Wno-unknown-pragmas --expt-relaxed-constexpr "$@" -c $f.cu -o ../../variants/obj_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc $ARCH -shared -o ../../variants/lib_$name.so ../../variants/obj_$name/*.o -lgomp
rm -rf ../../variants/obj_$name
echo built variants/lib_$name.so
