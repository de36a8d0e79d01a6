set -e
LIB=bayesianautoencoders_b200/lib/libbae_b200.so
cp $LIB /tmp/rel.so
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -DBAE_TC_DEBUG -o $LIB bayesianautoencoders_b200/csrc/bae_api.cu bayesianautoencoders_b200/csrc/bae_kernels.cu bayesianautoencoders_b200/csrc/bae_narrow.cu
for f in 0 1 2 4; do BAE_TC_DBGFLAGS=$f timeout 200 python tools/dev/exp_skipconv.py 2>&1 | tail -1; done
cp /tmp/rel.so $LIB
