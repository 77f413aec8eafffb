#!/bin/bash
# A/B builds of the main translation unit: tools/ab_build_api.sh <tag> [nvcc -D flags...] -> nps-waveform-analysis_b200/lib/libnpswf_<tag>.so
# (same ABI, the other objects of the regular build are reused; select with NPSWF_LIB=...)
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
P=nps-waveform-analysis_b200
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I include "$@" -c -o $P/lib/obj/npswf_api_$tag.o $P/csrc/npswf_api.cu
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $P/lib/libnpswf_$tag.so $P/lib/obj/npswf_api_$tag.o $P/lib/obj/npswf_migrad.o $P/lib/obj/host_pack.o $P/lib/obj/host_event.o
echo built $P/lib/libnpswf_$tag.so
