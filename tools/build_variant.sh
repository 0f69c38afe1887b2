#!/bin/bash
# Developer tool: build a library variant with extra -D flags for ONE translation unit (default: the reentry forward
# pass) into ssmtoybox_b200/lib/variants/libssmb200_<name>.so; select it at run time with SSM_B200_LIB=<path>.
#   tools/build_variant.sh <name> "<flags>" [source.cu]
set -e
name=$1; flags=$2; src=${3:-ssm_filter_reentry.cu}
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/build/variants $root/ssmtoybox_b200/lib/variants
obj=$root/build/variants/${name}_${src%.cu}.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -Xptxas -v $flags \
    -c $root/ssmtoybox_b200/csrc/$src -o $obj 2> $obj.ptxas.log
objs=$(ls $root/build/obj/*.o | grep -v "/${src%.cu}.o")
/usr/local/cuda/bin/nvcc -shared -o $root/ssmtoybox_b200/lib/variants/libssmb200_${name}.so $objs $obj -lcudart
grep -A3 "filter_pair_kernel" $obj.ptxas.log | grep "Used\|spill" | head -4
