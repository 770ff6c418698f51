#!/bin/bash
# experiment build of libdmx: tools/build_variant.sh NAME "-DDMX_X=1 ..."  ->  tools/libdmx_NAME.so (pick it with DMX_LIB=...)
set -e
HERE=$(cd $(dirname $0)/.. && pwd)
C=$HERE/demuxlet_b200/csrc
make -s -C $C
mkdir -p $HERE/tools/var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-ffp-contract=off -cudart static $2 -c $C/dmx_kernels.cu -o $HERE/tools/var/$1_kernels.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $HERE/tools/libdmx_$1.so $HERE/tools/var/$1_kernels.o $C/build/dmx_api.o $C/build/dmx_field.o $C/build/dmx_finalise.o $C/build/dmx_pack.o $C/build/dmx_vcf.o $C/build/dmx_bam.o -lpthread -lz
