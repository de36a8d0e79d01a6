#!/bin/bash
# Rebuild the library with ptxas statistics and report spill locations of the tcgen05 program kernel.
cd /root/repo/bayesianautoencoders_b200/csrc || exit 1
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -v -o ../lib/libbae_b200.so bae_api.cu bae_kernels.cu 2>&1 | grep -A2 "kernel_tc\|error" | grep -v "^--"
cuobjdump -sass ../lib/libbae_b200.so | awk '/Function : .*bae_program_kernel_tc/{f=1} f&&/Function : /&&!/bae_program_kernel_tc/{f=0} f' | grep "^        /\*[0-9a-f]*\*/" | sed 's/ *\/\* 0x[0-9a-f]* \*\///' > /tmp/prog_tc.sass
wc -l /tmp/prog_tc.sass
echo "spill instructions per 500-instruction bin:"
grep -n "LDL\|STL" /tmp/prog_tc.sass | awk -F: '{print int($1/500)*500}' | sort -n | uniq -c | head -12
grep -n "LDTM.x32\|USETMAXREG" /tmp/prog_tc.sass | head -4
