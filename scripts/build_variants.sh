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
    mb3) build mb3 -DWFB_V_MINBLOCKS=3 ;;
    mb5) build mb5 -DWFB_V_MINBLOCKS=5 ;;
    fused) build fused -DWFB_V_FUSED=1 ;;
    slowtrips) build slowtrips -DWFB_ENGINE_FAST_TRIPS=0 ;;
    a6) build a6 -DWFB_V_MINBLOCKS=6 ;;
    a8) build a8 -DWFB_V_MINBLOCKS=8 ;;
    c6) build c6 -DWFB_VC_MINBLOCKS=6 ;;
    c8) build c8 -DWFB_VC_MINBLOCKS=8 ;;
    a6c6) build a6c6 -DWFB_V_MINBLOCKS=6 -DWFB_VC_MINBLOCKS=6 ;;
    a4) build a4 -DWFB_V_MINBLOCKS=4 ;;
    c5) build c5 -DWFB_VC_MINBLOCKS=5 ;;
    c3) build c3 -DWFB_VC_MINBLOCKS=3 ;;
    pf0) build pf0 -DWFB_V_PREFETCH=0 ;;
    pf1) build pf1 -DWFB_V_PREFETCH=1 ;;
    pf1mb5) build pf1mb5 -DWFB_V_PREFETCH=1 -DWFB_V_MINBLOCKS=5 ;;
    pf2mb5) build pf2mb5 -DWFB_V_PREFETCH=2 -DWFB_V_MINBLOCKS=5 ;;
    pf2mb3) build pf2mb3 -DWFB_V_PREFETCH=2 -DWFB_V_MINBLOCKS=3 ;;
    mb6) build mb6 -DWFB_V_MINBLOCKS=6 ;;
    mb8) build mb8 -DWFB_V_MINBLOCKS=8 ;;
    mb5cs) build mb5cs -DWFB_V_MINBLOCKS=5 -DWFB_V_LDCS=1 ;;
    mb6cs) build mb6cs -DWFB_V_MINBLOCKS=6 -DWFB_V_LDCS=1 ;;
    mb4cs) build mb4cs -DWFB_V_MINBLOCKS=4 -DWFB_V_LDCS=1 ;;
    *) echo "unknown variant $v"; exit 1 ;;
  esac
done
ls -la _obj/variants/*.so
