#!/bin/bash
# Round 2: CNN variant check: cnn parity first (abort on failure), then full-size / trained round trips and the c2 bench line.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cnn" > gpurun_out/pytest_n0.log 2>&1 || { tail -5 gpurun_out/pytest_n0.log; echo "cnn tests failed"; exit 1; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cnn or full_size or trained" > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_n.log
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/n_c2.json 2> gpurun_out/n_c2.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/n_c2.json').read().strip().splitlines()[-1])
    print(round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'bpp', round(d['bpp'],4), 'cnn TF', round(d['cnn_tflops'],1))
except Exception as e: print('failed', e)
PY
