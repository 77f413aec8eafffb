#!/bin/bash
# A/B aid, runs on the GPU box: tools/ab_run.sh <base-tag> <tag>...  (libraries built by tools/ab_build_api.sh)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/ab_run.log
: > $out
for tag in "$@"; do
  NPSWF_LIB=$PWD/nps-waveform-analysis_b200/lib/libnpswf_$tag.so timeout 300 python tools/ab_variant.py $tag 3 >> $out 2>&1 || echo "variant $tag failed rc=$?" >> $out
done
base=$1; shift
python tools/ab_compare.py $base "$@" >> $out 2>&1
grep -E "^(AB|CMP|variant)" $out
