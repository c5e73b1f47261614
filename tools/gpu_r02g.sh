#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf -s > gpurun_out/pytest_r02g.log 2>&1; echo "pytest rc=$?"
grep -E "full-size|gradients, worst|negative control|passed|failed|FAILED" gpurun_out/pytest_r02g.log | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r02g.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r02g.json'))
print('mim', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'], 'roofline', j['roofline']['frac'], j['roofline']['launch_ms_by_shape'], j['roofline']['share_of_step'])
i=j['inference']; print('inf', i['value'], i['ms_per_step'], 'e2e', i['e2e']['value'], i['e2e_raw_int16']['value'], 'attn', i['roofline']['launch_ms'], i['roofline']['frac'])
print('cls', j['classification'].get('ms_per_step'), j['classification'].get('losses'))
print('vjepa', j['vjepa_step'].get('ms_per_step'), j['vjepa_encoder'].get('ms_per_volume'))
PY
timeout 200 python tools/run_cls.py 6 2>&1 | tail -3
