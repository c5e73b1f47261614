#!/bin/bash
# GEMM four-slab-buffer variant for the GELU(+pre) / dGELU epilogues: parity + sweep (on / off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_r03j.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r03j.log | cut -c1-300
for v in 1 0 1; do SMBV_GEMM_EPI4=$v python tools/gemm_small_sweep.py "SMBV_GEMM_EPI4=$v" 2>&1 | grep -E "EPI4|gelu|sum over"; done | tee gpurun_out/gemm_small_sweep_epi4.log
