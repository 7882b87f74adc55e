#!/bin/bash
# tools/build_variant.sh NAME FILE.cu [nvcc flags...]: rebuild one translation unit with extra flags
# and link it with the other objects into karios_b200/_lib/libkarios_b200_NAME.so (select it with
# KR_LIB=... for same-box A/B runs).  The default objects must exist (__graft_entry__.build()).
set -e
cd "$(dirname "$0")/.."
name=$1; file=$2; shift 2
obj=build/obj/${file%.cu}_$name.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
    -Xcompiler -fPIC,-fvisibility=hidden "$@" -c karios_b200/csrc/$file -o $obj
others=$(for o in build/obj/kr_api.o build/obj/kr_prep.o build/obj/kr_corners.o build/obj/kr_corner_fast.o build/obj/kr_sort.o build/obj/kr_lk.o build/obj/kr_zncc.o build/obj/kr_mi.o build/obj/kr_scene.o; do [ "$o" != "build/obj/${file%.cu}.o" ] && echo $o; done)
/usr/local/cuda/bin/nvcc -shared -o karios_b200/_lib/libkarios_b200_$name.so $obj $others
echo karios_b200/_lib/libkarios_b200_$name.so
