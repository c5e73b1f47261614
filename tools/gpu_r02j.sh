#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cuda_graph or data_parallel or three_opt" 2>&1 | tail -5
