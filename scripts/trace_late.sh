#!/bin/bash
# clock64 pipeline traces of gemm_tc on the late EfficientNet-B2 layer shapes (M = 64 clips x 64 positions)
mkdir -p gpurun_out
{
python scripts/gemm_check.py trace 4096 1248 208 0 1 0
python scripts/gemm_check.py trace 4096 208 1248 64 0 1
python scripts/gemm_check.py trace 16128 120 528 252 0 1
python scripts/gemm_check.py trace 16128 528 88 0 1 0
} > gpurun_out/trace_late.txt 2>&1
tail -5 gpurun_out/trace_late.txt
# fixed cost of a launch: one-tile GEMMs
python scripts/gemm_check.py trace 128 208 208 0 1 0 >> gpurun_out/trace_late.txt 2>&1
