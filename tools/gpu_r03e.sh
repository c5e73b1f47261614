#!/bin/bash
# programmatic dependent launch experiment (SMBV_PDL; measured no gain and REVERTED — the switch no longer exists in the library, see profiles/r02_attn_notes.md):
# parity subset, then the MIM step with PDL on / off (separate processes, same box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/pytest_r03e.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_r03e.log | cut -c1-300
for r in 1 2; do
  for pdl in 1 0; do
    echo "== SMBV_PDL=$pdl"
    SMBV_PDL=$pdl timeout 300 python tools/side_stream_ab.py 2 2>&1 | grep -E "round|on vs off" | tail -9
  done
done 2>&1 | tee gpurun_out/pdl_ab_r03e.log
