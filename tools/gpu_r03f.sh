#!/bin/bash
mkdir -p gpurun_out
for bn in 0 128 256; do SMBV_GEMM_BN=$bn python tools/gemm_small_sweep.py "SMBV_GEMM_BN=$bn"; done 2>&1 | tee gpurun_out/gemm_small_sweep.log
