#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r02d.sh
python tools/run_patch.py 20
timeout 900 python -m pytest tests/test_gpu_vjepa.py tests/test_gpu_examples.py -m gpu -q -rf > gpurun_out/pytest_r02e.log 2>&1; echo "pytest vjepa rc=$?"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_r02e.log | head -20
