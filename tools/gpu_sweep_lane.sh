#!/bin/bash
# Sweep of the lane schedule's threshold (warps of chains a band needs to take the lane kernel instead of the windowed schedule), c2.
mkdir -p gpurun_out
for w in 16 48 148 400; do
  LLICTI_LANE_MIN_WARPS=$w timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/sw_$w.json 2> gpurun_out/sw_$w.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/sw_$w.json').read().strip().splitlines()[-1])
    print('min warps $w:', 'value', round(d['value'],1), 'dec', round(d['decode_mpps']), 'decode ms', round(d['decode_ms_per_step'],2), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items() if k in ('decode','window','cnn','merge')})
except Exception as e: print('failed', e)
PY
done
