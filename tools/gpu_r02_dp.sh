#!/bin/bash
# DP tail experiment at N=2: NCCL CTA limits
mkdir -p gpurun_out
for cfg in "default" "NCCL_MAX_CTAS=2" "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=8" "NCCL_MAX_CTAS=4 NCCL_MIN_CTAS=1 NCCL_NVLS_ENABLE=0" ; do
  envs=""
  if [ "$cfg" != "default" ]; then envs="$cfg"; fi
  env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-inference --no-cls --no-vjepa 2>/dev/null | grep '^{"metric' > gpurun_out/dp_tmp.json
  python - "$cfg" <<'PY'
import json,sys
j=json.load(open('gpurun_out/dp_tmp.json'))
print(sys.argv[1], '| step', round(j['ms_per_step'],3), 'ms | no-collective', round(j['no_collective']['ms_per_step'],3), 'ms | e2e', round(j['e2e']['ms_per_step'],3))
PY
done
