#!/bin/bash
# Experimental builds of libglgym.so (fp64 nominal variants only): tools/ubench/build_exp.sh <name> <extra nvcc flags...>
# -> tools/ubench/exp_<name>.so (+ .sass); time it with tools/time_lib.py
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --fmad=true \
  -DGLG_DEV_FAST "$@" -shared -o tools/ubench/exp_$name.so greenlight-gym2_b200/csrc/glg_capi.cu -lcudart_static -ldl -lrt -lpthread \
  2> tools/ubench/exp_$name.log
grep -A2 "units_kernel" tools/ubench/exp_$name.log | grep -E "Used|spill" | paste - - | sed 's/ptxas info    : //g' | cut -c1-200
cuobjdump -sass tools/ubench/exp_$name.so > /tmp/w/exp_$name.sass
