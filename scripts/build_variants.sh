#!/bin/bash
# developer aid: build kernel-tuning variants of the library into wflow.jl_b200/csrc/_obj/variants/
set -e
cd "$(dirname "$0")/../wflow.jl_b200/csrc"
mkdir -p _obj/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2"
build() { # name, flags
  name=$1; shift
  d=_obj/variants/$name; mkdir -p $d
  for f in api vertical routing; do $NV "$@" -c $f.cu -o $d/$f.o & done
  $NV -x cu -c network.cpp -o $d/network.o &
  wait
  $NV -shared -o _obj/variants/lib_$name.so $d/api.o $d/vertical.o $d/routing.o $d/network.o
}
build ssf2 -DWFB_SSF_MINBLOCKS=2
build olf3 -DWFB_OLF_MINBLOCKS=3 -DWFB_RIV_MINBLOCKS=3
build v128 -DWFB_V_BLOCK=128 -DWFB_VA_MINBLOCKS=5 -DWFB_VC_MINBLOCKS=4
build v128b -DWFB_V_BLOCK=128 -DWFB_VA_MINBLOCKS=6 -DWFB_VC_MINBLOCKS=5
ls -la _obj/variants/*.so
