#!/bin/bash
# developer aid: build kernel-tuning variants of the library into wflow.jl_b200/csrc/_obj/variants/
# (select one with WFB_LIB=<path> python bench.py ...)
set -e
cd "$(dirname "$0")/../wflow.jl_b200/csrc"
mkdir -p _obj/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2"
build() { # name, flags
  name=$1; shift
  d=_obj/variants/$name; mkdir -p $d
  for f in api vertical routing local_inertial; do $NV "$@" -c $f.cu -o $d/$f.o & done
  $NV -x cu -c network.cpp -o $d/network.o &
  wait
  $NV -shared -o _obj/variants/lib_$name.so $d/api.o $d/vertical.o $d/routing.o $d/local_inertial.o $d/network.o
}
for v in "$@"; do
  case $v in
    slowtrips) build slowtrips -DWFB_ENGINE_FAST_TRIPS=0 ;;
    e5) build e5 -DWFB_ENGINE_MINBLOCKS=5 ;;
    e6) build e6 -DWFB_ENGINE_MINBLOCKS=6 ;;
    e8) build e8 -DWFB_ENGINE_MINBLOCKS=8 ;;
    a4) build a4 -DWFB_V_MINBLOCKS=4 ;;
    a6) build a6 -DWFB_V_MINBLOCKS=6 ;;
    c3) build c3 -DWFB_VC_MINBLOCKS=3 ;;
    c5) build c5 -DWFB_VC_MINBLOCKS=5 ;;
    a4c3) build a4c3 -DWFB_V_MINBLOCKS=4 -DWFB_VC_MINBLOCKS=3 ;;
    *) echo "unknown variant $v"; exit 1 ;;
  esac
done
ls -la _obj/variants/*.so
