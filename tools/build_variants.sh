#!/bin/bash
# tools/build_variants.sh "NT MINB [tag -Dflags...]" ... -- tuning builds of libjmb200 with other k_int_search launch shapes -> tools/_bin/
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_bin
for cfg in "$@"; do
  set -- $cfg
  out=tools/_bin/libjmb200_nt$1_b$2${3:+_$3}.so
  nt=$1; mb=$2; shift 2; [ $# -gt 0 ] && shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -DJMB_IS_NT=$nt -DJMB_IS_MINB=$mb $@ \
       -Xcompiler -fPIC -shared -cudart static -o $out jm_b200/csrc/jmb_context.cu jm_b200/csrc/k_subpel.cu jm_b200/csrc/k_search.cu jm_b200/csrc/k_refine.cu jm_b200/csrc/k_tq.cu jm_b200/csrc/k_epzs.cu jm_b200/csrc/k_chroma.cu jm_b200/csrc/k_deblock.cu
  echo built $out
done
