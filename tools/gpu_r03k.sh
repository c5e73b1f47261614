#!/bin/bash
mkdir -p gpurun_out
python tools/run_gemm_dgelu.py 20480 384 1536 dgelu
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 3 -c 1 -f -o gpurun_out/prof_r03k_dgelu python tools/run_gemm_dgelu.py 20480 384 1536 dgelu > gpurun_out/ncu_r03k.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_r03k.log
