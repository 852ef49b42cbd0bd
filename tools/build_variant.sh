#!/bin/bash
# build a variant of libssf.so with extra -D flags for A/B measurements:
#   tools/build_variant.sh <name> [-DFOO=1 ...]  ->  variants/libssf_<name>.so  (select with SSF_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p variants
C=supersurfel_fusion_b200/csrc
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -cudart static \
  --threads 4 "$@" -shared -o variants/libssf_$name.so $C/ssf_icp.cu $C/ssf_surfels.cu $C/ssf_tps.cu $C/ssf_ingest.cu $C/ssf_engine.cu
echo variants/libssf_$name.so
