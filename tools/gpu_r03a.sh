#!/bin/bash
# round-2, third session: second-stream weight-gradient queue — parity subset + A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "side_stream or grad or training_step or classification or cls" > gpurun_out/pytest_r03a.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_r03a.log | cut -c1-300
timeout 600 python tools/side_stream_ab.py 3 > gpurun_out/side_ab_r03a.log 2>&1; echo "ab rc=$?"
cat gpurun_out/side_ab_r03a.log | tail -30
