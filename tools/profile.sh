#!/bin/bash
# ncu evidence for bench.py (run under gpurun): launch list + one full capture of the dominant kernel.
# usage: tools/profile.sh <tag> [kernel-regex]
set -u
TAG=${1:-r01}
KREGEX=${2:-flash_attn_fwd}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mim"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s 12 -c 2 -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/plain_${TAG}.log
ls -la gpurun_out/
