#!/bin/bash
# N=2: overlapped per-bucket all-reduce vs one all-reduce after backward, CUDA-graph step (+ clean exit check)
mkdir -p gpurun_out
for cfg in "SMBV_DP_OVERLAP=1" "SMBV_DP_OVERLAP=0"; do
  env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-inference --no-cls --no-vjepa 2>/dev/null | grep '^{"metric' > gpurun_out/dp_tmp.json
  echo "rc=$? $cfg"
  python - "$cfg" <<'PY'
import json,sys
j=json.load(open('gpurun_out/dp_tmp.json'))
print(sys.argv[1], '| graph step', round(j['ms_per_step'],3), 'ms | eager', round(j['eager_ms_per_step'],3), '| no-collective', round(j['no_collective']['ms_per_step'],3), 'ms | e2e', round(j['e2e']['ms_per_step'],3))
PY
done
