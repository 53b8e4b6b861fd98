#!/bin/bash
# usage: tools/build_epzs_variants.sh "<int_minb> <sub_minb>" ...   -> tools/_bin/libjmb200_ei<int>_es<sub>.so (EPZS kernels at other occupancies)
set -e
mkdir -p tools/_bin
for cfg in "$@"; do
  set -- $cfg
  out=tools/_bin/libjmb200_ei$1_es$2.so
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -DJMB_EPZS_INT_MINB=$1 -DJMB_EPZS_SUB_MINB=$2 -diag-suppress 177 \
       -Xcompiler -fPIC -shared -cudart static -o $out jm_b200/csrc/jmb_context.cu jm_b200/csrc/k_subpel.cu jm_b200/csrc/k_search.cu jm_b200/csrc/k_refine.cu jm_b200/csrc/k_tq.cu jm_b200/csrc/k_epzs.cu jm_b200/csrc/k_chroma.cu jm_b200/csrc/k_deblock.cu
  echo built $out
done
