#!/bin/bash
tag=$1; mkdir -p gpurun_out
echo "== default" > gpurun_out/variants_$tag.log
timeout 300 python tools/f32_check.py venice >> gpurun_out/variants_$tag.log 2>&1
for so in sequencesfm_b200/variants/lib_*.so; do
  echo "== $so" >> gpurun_out/variants_$tag.log
  B200BA_LIB=$PWD/$so timeout 300 python tools/f32_check.py venice >> gpurun_out/variants_$tag.log 2>&1
done
grep "==\|cg_point_pass_f32\|fp32-PCG\|rror" gpurun_out/variants_$tag.log
