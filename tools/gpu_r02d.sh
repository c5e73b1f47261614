#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -rf -x -k "patch_embed or simmim or mim_forward or embedding or full_size or tiny" > gpurun_out/pytest_r02d.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_r02d.log
timeout 300 python tools/gpu_check.py patch > gpurun_out/gpu_check_r02d.log 2>&1; grep -E "patch" gpurun_out/gpu_check_r02d.log | cut -c1-300
