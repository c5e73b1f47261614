#!/bin/bash
# NOTE for the next session: the ncu launch list of the bench command below hit its 900 s limit (the bench now runs three passes of the MIM step;
# ncu serialises ~6000 launches) and cost 15 GPU-minutes: capture `python tools/run_train.py 1` instead (one step, tools/summarize_train_launches.py).
# third session, evidence call: one ncu --set full launch of every hot kernel at its benchmark shape (final build) + the ncu launch list of the bench command
mkdir -p gpurun_out
timeout 280 python tools/ncu_targets.py > gpurun_out/ncu_targets_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:'flash_attn_fwd2|flash_attn_bwd_|gemm_bf16|patch_embed|normpix_loss|layernorm_fwd|rope3d|adamw' \
  -o gpurun_out/prof_r03n_kernels -f python tools/ncu_targets.py > gpurun_out/ncu_r03n.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_targets_plain.log; tail -3 gpurun_out/ncu_r03n.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cls --no-vjepa > gpurun_out/bench_r03n_plain.json 2> gpurun_out/bench_r03n_plain.err; echo "bench plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r03n.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cls --no-vjepa > gpurun_out/ncu_bench_r03n.log 2>&1; echo "ncu launch list rc=$?"
ls -la gpurun_out/launches_bench_r03n.csv gpurun_out/prof_r03n_kernels.ncu-rep
