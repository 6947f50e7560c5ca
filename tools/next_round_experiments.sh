#!/bin/bash
# Run on the GPU box (one gpurun call, ~3 minutes): A/B timings of the kernel variants prepared but not yet measured.
# Every variant is compile-time only (-D), built into gpurun_out/, the shipped library is untouched; see DESIGN.md
# section 8.  Numbers are us per launch at C2 (256 x 1 s) and at 2048 clips.
set -x
python tools/variant_bench.py "" \
    "-DB200MEL_MEL_FIXED=0x731" \
    "-DB200MEL_WARPS_PER_CTA=12" \
    "-DB200MEL_WARPS_PER_CTA=12 -DB200MEL_MEL_FIXED=0x731" \
    "-DB200MEL_X_NOMEL"
VB_CLIPS=2048 python tools/variant_bench.py "" "-DB200MEL_MEL_FIXED=0x731" "-DB200MEL_WARPS_PER_CTA=12 -DB200MEL_MEL_FIXED=0x731"
VB_WORKLOAD=C4 python tools/variant_bench.py "" "-DB200MEL_SPLIT_LDS64"
VB_WORKLOAD=C5 python tools/variant_bench.py "" "-DB200MEL_MEL_FIXED=0xa32"
rm -f gpurun_out/*.so gpurun_out/*.pt
