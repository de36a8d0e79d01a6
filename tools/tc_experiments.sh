#!/bin/bash
# Timing experiments on the tcgen05 engine. The debug-flag branches exist only in a -DBAE_TC_DEBUG build (results are
# WRONG when a flag is set; timing only), so this script builds its own library next to the shipped one and restores
# the production build afterwards.
set -e
cd "$(dirname "$0")/.."
LIB=bayesianautoencoders_b200/lib/libbae_b200.so
cp $LIB /tmp/libbae_b200.release.so
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DBAE_TC_DEBUG \
  -o $LIB bayesianautoencoders_b200/csrc/bae_api.cu bayesianautoencoders_b200/csrc/bae_kernels.cu
for f in ${FLAGS:-0 1 8 16 17 25}; do
  echo "== BAE_TC_DBGFLAGS=$f"
  BAE_TC_DBGFLAGS=$f timeout 120 python tools/dev/dev_trace.py 2>&1 | tail -12
done
cp /tmp/libbae_b200.release.so $LIB
