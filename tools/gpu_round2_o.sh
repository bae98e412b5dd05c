#!/bin/bash
# Round 2: pipelined host entry points + incremental lane positions: parity, c2 bench (value and e2e).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or lane or round_trip or full_size or mixed or malformed" > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_o.log
for mode in default "0"; do
  if [ "$mode" = "default" ]; then unset LLICTI_HOST_PIPELINE; else export LLICTI_HOST_PIPELINE=$mode; fi
  timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/o_c2_$mode.json 2> gpurun_out/o_c2_$mode.err
  echo "bench $mode rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/o_c2_$mode.json').read().strip().splitlines()[-1])
    print('$mode', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'e2e enc', round(d['e2e']['encode_mpps']), 'e2e dec', round(d['e2e']['decode_mpps']), 'dec kernel', round(d['kernel_ms_per_step']['decode'],2))
except Exception as e: print('failed', e)
PY
done
