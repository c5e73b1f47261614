#!/bin/bash
# round-2 GPU call A: tests, bench, pipe-rate microbenchmark, kernel timings, one ncu --set full pass over every hot kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_a.log 2>&1
nproc >> gpurun_out/smi_a.log
timeout 1200 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_r02a.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_r02a.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r02a.err
timeout 120 tools/microbench/pipe_rates > gpurun_out/pipe_rates.log 2>&1; echo "pipe_rates rc=$?"
cat gpurun_out/pipe_rates.log
timeout 400 python tools/gpu_check.py vjepa gemm patch > gpurun_out/gpu_check_r02a.log 2>&1; echo "gpu_check rc=$?"
grep -E "rope|time" gpurun_out/gpu_check_r02a.log | cut -c1-300
timeout 280 python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'flash_attn_fwd2|flash_attn_bwd_d|gemm_bf16|patch_embed|normpix_loss|layernorm_fwd|rope3d|adamw' \
  -o gpurun_out/prof_r02_kernels -f python tools/ncu_targets.py > gpurun_out/ncu_r02.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_targets_plain.log; tail -5 gpurun_out/ncu_r02.log
