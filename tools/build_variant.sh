#!/bin/bash
# tools/build_variant.sh NAME [-DFLAG ...]  -> build/variants/NAME.so (experiments only; the product library is csrc/libclothgpu.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xptxas -v \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden -ccbin /usr/bin/g++ -shared \
  -o build/variants/$name.so cloth-physics_b200/csrc/clothgpu_api.cu -lcudart "$@" 2> build/variants/$name.ptxas.log
grep -A2 "${PTXAS_KERNEL:-wave_frame_kernelIdLb0ELb0}" build/variants/$name.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo " [$name]"
