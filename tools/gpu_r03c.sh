#!/bin/bash
# N=2 (third session): data-parallel MIM step with the second-stream queue (bucket hook on the side stream, early fold-back) vs the
# single-stream step (SMBV_WGRAD_STREAM=0), CUDA-graph step, same box; then the replica-consistency check
mkdir -p gpurun_out
for cfg in "SMBV_WGRAD_STREAM=1" "SMBV_WGRAD_STREAM=0" "SMBV_WGRAD_STREAM=1"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-inference --no-cls --no-vjepa --no-cpu-baseline 2>gpurun_out/dp_r03c.err | grep '^{"metric' > gpurun_out/dp_tmp.json
  echo "rc=$? $cfg"; tail -c 300 gpurun_out/dp_r03c.err
  python - "$cfg" <<'PY'
import json,sys
j=json.load(open('gpurun_out/dp_tmp.json'))
print(sys.argv[1], '| graph step', round(j['ms_per_step'],3), 'ms | eager', round(j['eager_ms_per_step'],3), '| no-collective', round(j['no_collective']['ms_per_step'],3), 'ms | e2e', round(j['e2e']['ms_per_step'],3), '| loss', j.get('loss_last'))
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_dp.py 2>&1 | tail -6
