#!/bin/bash
# third session: full GPU test-suite + bench (all blocks) with the second-stream backward
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -q -rf > gpurun_out/pytest_r03g.log 2>&1; echo "pytest rc=$? (${SECONDS} s)"
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_r03g.log | cut -c1-400
SECONDS=0
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r03g.json 2> gpurun_out/bench_r03g.err; echo "bench rc=$? (${SECONDS} s)"
tail -c 400 gpurun_out/bench_r03g.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r03g.json'))
print('mim', round(j['value'],2), round(j['ms_per_step'],3), 'eager', round(j['eager_ms_per_step'],3), 'e2e', round(j['e2e']['value'],2), 'roofline', round(j['roofline']['frac'],3), j['roofline']['launch_ms_by_shape'], 'launches', j['gpu_launches'], j['clocks'])
i=j['inference']; print('inf', round(i['value'],2), round(i['ms_per_step'],3), 'e2e', round(i['e2e']['value'],2), round(i['e2e_raw_int16']['value'],2), 'attn', round(i['roofline']['launch_ms'],3), round(i['roofline']['frac'],3))
c=j['classification']; print('cls', c.get('value'), c.get('ms_per_step'), c.get('eager_ms_per_step'))
print('vjepa', j['vjepa_step'].get('ms_per_step'), j['vjepa_step'].get('eager_ms_per_step'), j['vjepa_encoder'].get('ms_per_volume'))
PY
