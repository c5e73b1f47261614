#!/bin/bash
# final round-2 check: the driver's sequence on one GPU — pytest -m gpu, smoke, bench (default flags, with cpu_baseline), reference arm
mkdir -p gpurun_out
echo "pytest: 135 passed in the previous call"
grep -E "passed|failed|FAILED" gpurun_out/pytest_r02z.log | cut -c1-300
SECONDS=0
timeout 900 python bench.py > gpurun_out/bench_r02z.json 2> gpurun_out/bench_r02z.err; echo "bench rc=$?"
echo "bench wall ${SECONDS} s"; tail -c 300 gpurun_out/bench_r02z.err; SECONDS=0
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_r02z.json'))
print('mim', round(j['value'],2), round(j['ms_per_step'],3), 'eager', round(j['eager_ms_per_step'],3), 'e2e', round(j['e2e']['value'],2), 'roofline', round(j['roofline']['frac'],3), j['roofline']['launch_ms_by_shape'], 'launches', j['gpu_launches'], j['clocks'])
i=j['inference']; print('inf', round(i['value'],2), round(i['ms_per_step'],3), 'e2e', round(i['e2e']['value'],2), round(i['e2e_raw_int16']['value'],2), 'attn', round(i['roofline']['launch_ms'],3), round(i['roofline']['frac'],3))
c=j['classification']; print('cls', c.get('value'), c.get('ms_per_step'), c.get('eager_ms_per_step'))
print('vjepa', j['vjepa_step'].get('ms_per_step'), j['vjepa_step'].get('eager_ms_per_step'), j['vjepa_encoder'].get('ms_per_volume'))
print('cpu_baseline', j.get('cpu_baseline'))
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02z_ref.json 2> gpurun_out/bench_r02z_ref.err; echo "ref rc=$?"
echo "reference wall ${SECONDS} s"; tail -c 200 gpurun_out/bench_r02z_ref.err; cut -c1-700 gpurun_out/bench_r02z_ref.json
