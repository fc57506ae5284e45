#!/bin/bash
# Experiment builds of libpoutine_b200.so with extra -D flags:  tools/build_variant.sh NAME "-DPB_KX_STAGES=8 ..."
# -> exp/libpb_NAME.so (git-ignored, travels with gpurun); run with PB_LIB=exp/libpb_NAME.so python bench.py
set -e
cd "$(dirname "$0")/../poutine_b200/csrc"
name=$1; flags=$2
out=../../exp; mkdir -p $out; tmp=$(mktemp -d)
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags"
for f in pb_api.cu k1_events.cu k3_null.cu pb_comm.cu pb_group.cpp nccl_dyn.cpp packer.cpp fasta.cpp; do $NV -c $f -o $tmp/${f%.*}.o & done
$NV --fmad=false -c k4_pvalues.cu -o $tmp/k4_pvalues.o &
$NV --fmad=false -c k5_extras.cu -o $tmp/k5_extras.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpb_$name.so $tmp/*.o -lcudart -ldl -lpthread
rm -rf $tmp; echo built exp/libpb_$name.so
