#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_small_sweep.py three-stages 2>&1 | tee gpurun_out/gemm_small_sweep_3s.log
timeout 300 python tools/gpu_check.py gemm 2>&1 | grep -E "gemm_time|FAIL" | cut -c1-250
