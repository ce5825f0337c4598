#!/bin/bash
# Builds a VARIANT of librt_b200.so with extra nvcc flags (A/B experiments, diagnostics builds) into
# cs420-ray-tracer_b200/build/<name>/librt_b200.so; use it with RTB200_LIB=<that path>.   usage: build_variant.sh name -DFLAG ...
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)
P=$HERE/cs420-ray-tracer_b200
name=$1; shift
O=$P/build/$name; mkdir -p $O
NV="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-Wall -I $HERE/include -I $P/csrc"
nvcc $NV "$@" -Xptxas -v -c $P/csrc/rt_kernels.cu -o $O/rt_kernels.o 2> $O/rt_kernels.ptxas.log &
nvcc $NV "$@" -c $P/csrc/rt_api.cu -o $O/rt_api.o &
nvcc $NV "$@" -x cu -c $P/csrc/scene_io.cpp -o $O/scene_io.o &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $O/librt_b200.so $O/rt_api.o $O/rt_kernels.o $O/scene_io.o -cudart static
echo $O/librt_b200.so
