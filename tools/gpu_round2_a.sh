#!/bin/bash
# first GPU call of round 2: GPU tests, default bench, group-size sweep, oracle self-check
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for g in 2 4 8 16; do
  LLICTI_GROUP_LANES=$g timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/bench_c2_g$g.json 2> gpurun_out/bench_c2_g$g.err
  echo "G=$g rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c2_g$g.json').read().strip().splitlines()[-1])
    print('G=$g', round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['decode_stats_per_step']['slow_path_symbols'])
except Exception as e: print('G=$g failed', e)
PY
done
LLICTI_GROUP_LANES=4 timeout 300 python bench.py --workload c2 --decode-impl 2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/bench_c2_windows.json 2> gpurun_out/bench_c2_windows.err; echo "windows rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "default bench rc=$?"
tail -c 600 gpurun_out/bench_default.err
timeout 400 python tools/oracle_selfcheck.py --reps 1 > gpurun_out/oracle_selfcheck.log 2>&1; echo "selfcheck rc=$?"
tail -8 gpurun_out/oracle_selfcheck.log
