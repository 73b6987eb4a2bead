#!/bin/bash
# Time the per-kernel groups for the default library and every variant library.  Usage: tools/gpu_variants.sh <tag>
tag=$1; mkdir -p gpurun_out
echo "== default" > gpurun_out/variants_$tag.log
timeout 300 python tools/prof_cg.py venice >> gpurun_out/variants_$tag.log 2>&1
for so in sequencesfm_b200/variants/lib_*.so; do
  echo "== $so" >> gpurun_out/variants_$tag.log
  B200BA_LIB=$PWD/$so timeout 300 python tools/prof_cg.py venice >> gpurun_out/variants_$tag.log 2>&1
done
grep "==\|cg_point_pass \|cg_camera_pass \|cg_step\|error\|Error" gpurun_out/variants_$tag.log
