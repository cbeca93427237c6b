#!/bin/bash
# tools/ab_build.sh NAME [nvcc -D flags...]  -> build/exp/libpb_NAME.so : an experiment build of the aligner for A/B runs
# on the GPU box (PB_LIB=build/exp/libpb_NAME.so python tools/profile_step.py ...).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/exp
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden "$@" \
     -c pacbioassembly_b200/csrc/pb_align.cu -o build/exp/pb_align_$name.o
nvcc -shared -o build/exp/libpb_$name.so build/obj/pb_ctx.o build/obj/pb_seq.o build/obj/pb_seed.o build/obj/pb_alignw.o build/obj/pb_locate.o build/obj/pb_pairs.o build/obj/pb_cons.o \
     build/exp/pb_align_$name.o -gencode arch=compute_100a,code=sm_100a
echo build/exp/libpb_$name.so
