#!/bin/bash
# Developer builds of the SAME ABI for A/B timing on one box (loaded through SUNB200_LIB; never shipped, *.so is git-ignored):
#   tools/_ab/libsunb200_nopdl.so   -DSUNB_NO_PDL            (tools/ab_pdl.sh)
#   tools/_ab/libsunb200_trace.so   -DSUNB_TAIL_TRACE        (`tools/build_variants.sh trace`; tools/tail_trace.py prints the event timeline
#                                                             of one work item of the fused block tail)
#   tools/_dbg/libsunb_dbg<v>.so    -DSUNB_TAIL_DBG=<v>      (tools/tail_exp.sh: the fused block tail with parts compiled out;
#                                                             bits: 1 no conv3, 2 one tap, 4 no GELU, 8 no stores, 16 / 32 no loads)
set -e
cd "$(dirname "$0")/../few-shot-vit_b200/csrc"
SRCS=$(sed -n 's/^SRCS *:= *//p' Makefile)
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
build() {   # build <output .so> <extra flags>
  local out=$1; shift
  local tmp=$(mktemp -d)
  for f in $SRCS; do nvcc $FLAGS "$@" -c $f -o $tmp/${f%.cu}.o & done; wait
  mkdir -p "$(dirname "$out")"
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$out" $tmp/*.o
  rm -rf $tmp
  echo "built $out"
}
build ../../tools/_ab/libsunb200_nopdl.so -DSUNB_NO_PDL
if [ "$1" = "trace" ]; then build ../../tools/_ab/libsunb200_trace.so -DSUNB_TAIL_TRACE; fi
if [ "$1" = "tail" ]; then
  for v in 1 2 4 8 63; do build ../../tools/_dbg/libsunb_dbg$v.so -DSUNB_TAIL_DBG=$v; done
fi
