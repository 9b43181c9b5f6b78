#!/bin/bash
# Builds libfgq_cuda.so into a scratch location and swaps it into the package in one rename, so that a snapshot of
# the tree (gpurun) never sees a half-written library.
R=/root/repo
mkdir -p /tmp/fgqbuild
make -C $R/fastgaussquadrature.jl_b200/csrc -j8 OUT=/tmp/fgqbuild/libfgq_cuda.so > /tmp/fgqbuild/make.log 2>&1
rc=$?
grep -v "^/usr/local/cuda/bin/nvcc\|^make" /tmp/fgqbuild/make.log | tail -20
if [ $rc -ne 0 ]; then echo "BUILD FAILED"; exit 1; fi
if [ "$1" = "noswap" ]; then echo "built (not swapped in: /tmp/fgqbuild/libfgq_cuda.so)"; exit 0; fi
cp /tmp/fgqbuild/libfgq_cuda.so $R/fastgaussquadrature.jl_b200/libfgq_cuda.so.new
mv $R/fastgaussquadrature.jl_b200/libfgq_cuda.so.new $R/fastgaussquadrature.jl_b200/libfgq_cuda.so
echo built
