#!/bin/bash
# N=8 check of the bench command the driver runs (CUDA-graph step with the captured all-reduce), bounded
mkdir -p gpurun_out
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-vjepa > gpurun_out/bench_r02m_n8.json 2> gpurun_out/bench_r02m_n8.err
echo "rc=$?"
tail -c 600 gpurun_out/bench_r02m_n8.err
grep '^{"metric' gpurun_out/bench_r02m_n8.json | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('N8 mim', j['value'], j['ms_per_step'], 'eager', j['eager_ms_per_step'], 'e2e', j['e2e']['value'], 'nocoll', j['no_collective']['ms_per_step'])
i=j.get('inference',{}); print('inf', i.get('value'), i.get('ms_per_step'), i.get('e2e',{}).get('value'))
c=j.get('classification',{}); print('cls', c.get('value'), c.get('ms_per_step'), c.get('error'))
"
