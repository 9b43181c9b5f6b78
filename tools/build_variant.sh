#!/bin/bash
# build_variant.sh <name> <extra nvcc flags...>: an experimental build of the library with legendre.cu recompiled
# under the given flags, into _exp/libfgq_<name>.so (timed with tools/time_tail.py via FGQ_LIB_PATH)
name=$1; shift
C=/root/repo/fastgaussquadrature.jl_b200/csrc
mkdir -p /root/repo/_exp
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC -Xcompiler -ffp-contract=off -Xptxas -v "$@" -c $C/legendre.cu -o /tmp/legendre_$name.o 2> /tmp/leg_$name.log || { tail -20 /tmp/leg_$name.log; exit 1; }
grep -A2 "legendre_asy_kernelILi0ELi0ELb0" /tmp/leg_$name.log | tail -2
objs=$(ls $C/_build/*.o | grep -v legendre.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/_exp/libfgq_$name.so $objs /tmp/legendre_$name.o -lcudart -ldl
