#!/bin/bash
# A/B of the lane decoder's occupancy target (LLICTI_LANE_OCC = CTAs per SM the register allocation aims at) on the c2 workload.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane or full_size or trained" > gpurun_out/pytest_lane.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_lane.log
for occ in 4 5 6; do
  LLICTI_LANE_OCC=$occ timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-per-config > gpurun_out/occ_$occ.json 2> gpurun_out/occ_$occ.err; echo "occ $occ rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/occ_$occ.json').read().strip().splitlines()[-1])
    print('occ $occ value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e: print('failed', e)
PY
done
LLICTI_LANE_OCC=6 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-per-config --weights-npz tests/golden/ckpt_A_trained.npz > gpurun_out/occ_6_tw.json 2> gpurun_out/occ_6_tw.err; echo "occ 6 trained rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/occ_6_tw.json').read().strip().splitlines()[-1])
    print('occ 6 trained value', round(d['value'],1), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e: print('failed', e)
PY
