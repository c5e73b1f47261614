#!/bin/bash
# GEMM with four epilogue groups: parity (GPU tests touching the GEMMs) + the small-shape sweep + full-size shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_r03h.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r03h.log | cut -c1-300
python tools/gemm_small_sweep.py four-groups 2>&1 | tee gpurun_out/gemm_small_sweep_4g.log
timeout 300 python tools/gpu_check.py gemm 2>&1 | grep -E "gemm_time|FAIL" | cut -c1-250
