#!/bin/bash
# A/B variants of the CUDA library from the working tree: build_variants/lib_<name>.so with extra -D flags.
#   tools/build_variants.sh name1 "-DCTC_OPT_CH=1" name2 "-DCTC_OPT_CH=1 -DCTC_OPT_FAST2=0" ...
# (bench with CTCDECODE_B200_LIB=$PWD/build_variants/lib_<name>.so; kernels vary by a few % between GPU boxes of the pool,
#  so variants are only comparable inside ONE gpurun call)
mkdir -p build_variants
while [ $# -ge 2 ]; do
  n=$1; f=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 1886 $f \
      -c ctcdecode_b200/csrc/ctc_api.cu -o build_variants/obj_$n.o 2>&1 | grep -E "error" ; \
    g++ -shared -nostdlib++ -o build_variants/lib_$n.so build_variants/obj_$n.o -L/usr/local/cuda/lib64 -lcudart_static -l:libstdc++.so.6 -lm -lrt -lpthread -ldl && rm -f build_variants/obj_$n.o && echo built $n ) &
  if [ $(jobs -r | wc -l) -ge 4 ]; then wait -n; fi
done
wait
