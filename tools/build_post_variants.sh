#!/bin/bash
# Build variants of the post-BA stage kernels (compile-time knobs POST_HG / POST_PPT / POST_MINB) for one-call A/B timing
# on the GPU box: tools/build_post_variants.sh name:"-DPOST_PPT=1" ...  -> sequencesfm_b200/variants/lib_<name>.so
# (the BA engine object is reused from build/obj); run each with B200BA_LIB=... python tools/post_bench.py --no-cpu
set -e
cd "$(dirname "$0")/.."
make -s -C sequencesfm_b200/csrc >/dev/null 2>&1
mkdir -p sequencesfm_b200/variants
rm -f sequencesfm_b200/variants/*.so
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-ffp-contract=off --compiler-bindir /usr/bin/g++ -Xptxas -v $flags \
     -c -o /tmp/post_$name.o sequencesfm_b200/csrc/post_stage.cu 2> /tmp/post_variant_$name.log || { echo "BUILD FAILED $name"; tail -5 /tmp/post_variant_$name.log; continue; }
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o sequencesfm_b200/variants/lib_$name.so build/obj/ba_engine.o /tmp/post_$name.o -ldl
  echo "$name [$flags]: $(grep -A2 'k_post_histILb0\|k_post_score\|k_post_gather' /tmp/post_variant_$name.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
done
