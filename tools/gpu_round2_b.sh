#!/bin/bash
# second GPU call: group decoder with parameter prefetch -- per-launch times, group-size sweep, ncu of the scale-0 launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "narrow or group or forward or fingerprint or offsets or contexts or round_trip_lossless or starved or schedules_agree or graph_replay" > gpurun_out/pytest_b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_b.log
for g in 4 8 16; do
  LLICTI_GROUP_LANES=$g timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/b_c2_g$g.json 2> gpurun_out/b_c2_g$g.err
  echo "G=$g rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/b_c2_g$g.json').read().strip().splitlines()[-1])
    print('G=$g', round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['decode_stats_per_step']['slow_path_symbols'])
except Exception as e: print('G=$g failed', e)
PY
done
LLICTI_PROF_DUMP=1 LLICTI_GROUP_LANES=8 timeout 300 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config > gpurun_out/b_c2_dump.json 2> gpurun_out/b_c2_dump.err
grep "class 6" gpurun_out/b_c2_dump.err | head -40
CMD="python bench.py --workload c2 --images 4 --steps 1 --warmup 1 --no-cpu --no-per-config"
LLICTI_GROUP_LANES=8 $CMD > gpurun_out/plain_ncu.log 2>&1 && LLICTI_GROUP_LANES=8 ncu --set full --clock-control none --import-source on -k regex:decode_band_group -s 12 -c 3 -o gpurun_out/r02_group $CMD > gpurun_out/ncu_group.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_group.log
