#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "round_trip or group or full_size or kodak or trained" > gpurun_out/pytest_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_e.log
for cfg in "G4 LLICTI_GROUP_LANES=4" "G2 LLICTI_GROUP_LANES=2" "G4w300 LLICTI_GROUP_LANES=4 LLICTI_GROUP_MIN_WARPS=300"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/e_c2_$name.json 2> gpurun_out/e_c2_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/e_c2_$name.json').read().strip().splitlines()[-1])
    print('$name', round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['decode_stats_per_step']['slow_path_symbols'], 'bpp', round(d['bpp'],4), d['bpp_delta_vs_reference_streams']['value'])
except Exception as e: print('$name failed', e)
PY
done
LLICTI_PROF_DUMP=1 timeout 300 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config > gpurun_out/e_c2_dump.json 2> gpurun_out/e_c2_dump.err
grep -E "class (6|8|1) " gpurun_out/e_c2_dump.err | tail -75 | awk '{printf "%s:%s ", $6, $7} END {print ""}'
CMD="python bench.py --workload c2 --images 6 --steps 1 --warmup 1 --no-cpu --no-per-config"
export LLICTI_GROUP_MIN_WARPS=1
$CMD > gpurun_out/plain_ncu_e.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:decode_band_group -s 12 -c 1 -o gpurun_out/r02_group_locate $CMD > gpurun_out/ncu_group_e.log 2>&1; echo "ncu rc=$?"
