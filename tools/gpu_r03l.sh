#!/bin/bash
# polynomial erf (no MUFU) in the GELU / dGELU epilogues + EPI4 policy: parity + sweep + full-size GEMM timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_r03l.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r03l.log | cut -c1-300
python tools/gemm_small_sweep.py "poly-erf" 2>&1 | tee gpurun_out/gemm_small_sweep_poly.log | grep -E "gelu|sum over"
timeout 300 python tools/gpu_check.py gemm 2>&1 | grep -E "gemm_time|gelu|FAIL" | cut -c1-250
