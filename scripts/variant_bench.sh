#!/bin/bash
# Build variants of the kernels (only W=16384) with different -D flags on the GPU box and time them.
# env: ONLY_LOG2H (log2(W)-1 to compile, default 13), KT_ARGS (arguments of kernel_time.py)
# usage: bash scripts/variant_bench.sh "name1:-DFOO=1 -DBAR=2" "name2:..." ...
mkdir -p gpurun_out/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  out=gpurun_out/variants/lib_${name}.so
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -shared \
       -DRCG_ONLY_LOG2H=${ONLY_LOG2H:-13} $flags -o $out rocoder_b200/csrc/*.cu -ldl 2> gpurun_out/variants/${name}.build.log \
    || { echo "{\"tag\": \"$name\", \"error\": \"build failed\"}"; continue; }
  RCG_LIB=$out python scripts/kernel_time.py --tag "$name" $KT_ARGS 2>> gpurun_out/variants/${name}.build.log
done
