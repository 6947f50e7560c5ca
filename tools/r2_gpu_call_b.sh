#!/bin/bash
# Round 2, GPU call B: the whole GPU test suite (new full-shape / round-2 tests included), then A/B timings of the
# fast-path kernel variants, then a bench line.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python tools/variant_bench.py "" "ENV:B200MEL_NO_FAST=1" "-DB200MEL_FAST_MEL_HOIST" "-DB200MEL_FAST_WIN_TABLE" \
    "-DB200MEL_FAST_WIN_TABLE -DB200MEL_FAST_MEL_HOIST" "-DB200MEL_WARPS_PER_CTA=12" "-DB200MEL_WARPS_PER_CTA=14"
VB_CLIPS=2048 python tools/variant_bench.py "" "ENV:B200MEL_NO_FAST=1" "-DB200MEL_FAST_WIN_TABLE"
VB_WORKLOAD=C5 python tools/variant_bench.py "" "ENV:B200MEL_NO_FAST=1" "-DB200MEL_FAST_WIN_TABLE"
VB_WORKLOAD=C4 python tools/variant_bench.py "" "-DB200MEL_SPLIT_LDS64"
python bench.py --steps 200 --warmup 10 | tee gpurun_out/r2_bench_b.json
rm -f gpurun_out/*.so gpurun_out/*.pt
