#!/bin/bash
mkdir -p gpurun_out
run() { # name, env..., then bench args after --
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
for cfg in "G8 LLICTI_GROUP_LANES=8" "G16 LLICTI_GROUP_LANES=16" "G8w1184 LLICTI_GROUP_LANES=8 LLICTI_GROUP_MIN_WARPS=1184" "G4 LLICTI_GROUP_LANES=4"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/c_c2_$name.json 2> gpurun_out/c_c2_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c_c2_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['decode_stats_per_step']['slow_path_symbols'])
except Exception as e: print('$name failed', e)
PY
done
LLICTI_PROF_DUMP=1 timeout 300 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config > gpurun_out/c_c2_dump.json 2> gpurun_out/c_c2_dump.err
grep -E "class (6|8|1) " gpurun_out/c_c2_dump.err | tail -75 | awk '{printf "%s:%s ", $6, $7} END {print ""}'
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_c.log
