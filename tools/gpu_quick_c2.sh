#!/bin/bash
# Quick check of a kernel change: the tests named in $1 (pytest -k expression), then the c2 bench line (no CPU leg, no per_config).
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$1" > gpurun_out/pytest_q.log 2>&1 || { tail -8 gpurun_out/pytest_q.log; echo "tests failed"; exit 1; }
tail -1 gpurun_out/pytest_q.log
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/q_c2.json 2> gpurun_out/q_c2.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/q_c2.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'cnn TF', round(d['cnn_tflops'],1), 'newton', d['decode_stats_per_step']['consumer_polls'], 'extra', d['decode_stats_per_step']['slow_path_symbols'])
except Exception as e: print('failed', e)
PY
