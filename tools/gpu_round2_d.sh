#!/bin/bash
mkdir -p gpurun_out
for cfg in "G4 LLICTI_GROUP_LANES=4" "G2 LLICTI_GROUP_LANES=2" "G8 LLICTI_GROUP_LANES=8" "G8nolocate LLICTI_GROUP_LANES=8 LLICTI_GROUP_LOCATE=0" "G4bf16 LLICTI_GROUP_LANES=4 LLICTI_TC_OPERANDS=bf16"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/d_c2_$name.json 2> gpurun_out/d_c2_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/d_c2_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['decode_stats_per_step']['slow_path_symbols'], 'bpp', round(d['bpp'],4))
except Exception as e: print('$name failed', e)
PY
done
LLICTI_PROF_DUMP=1 timeout 300 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config > gpurun_out/d_c2_dump.json 2> gpurun_out/d_c2_dump.err
grep -E "class (6|8|1) " gpurun_out/d_c2_dump.err | tail -75 | awk '{printf "%s:%s ", $6, $7} END {print ""}'
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_d.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_d.log
