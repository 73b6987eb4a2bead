#!/bin/bash
# Build kernel-variant libraries (compile-time knobs) for one-call A/B timing on the GPU box.
# Usage: tools/build_variants.sh name1:"-DX=1 -DY=2" name2:"..." ; outputs sequencesfm_b200/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p sequencesfm_b200/variants
rm -f sequencesfm_b200/variants/*.so
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --compiler-bindir /usr/bin/g++ -Xptxas -v $flags \
     -shared -o sequencesfm_b200/variants/lib_$name.so sequencesfm_b200/csrc/ba_engine.cu -ldl 2> /tmp/variant_$name.log || { echo "BUILD FAILED $name"; tail -5 /tmp/variant_$name.log; continue; }
  echo "$name [$flags]: $(grep -A2 'k_cg_camera_passENS' /tmp/variant_$name.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ') | tab: $(grep -A2 'k_cg_point_pass_tabENS' /tmp/variant_$name.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
done
