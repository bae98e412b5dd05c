#!/bin/bash
# Round 2, call h: CNN with the hidden activations in TMEM + lane decoder with cheap prefetch addressing: full GPU suite, bench, ncu.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_h.log
for cfg in "lane LLICTI_DECODE_LANES=1" "lane_w600 LLICTI_LANE_MIN_WARPS=600" "lane_w300 LLICTI_LANE_MIN_WARPS=300"; do
  set -- $cfg; name=$1; shift
  env "$@" timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/h_c2_$name.json 2> gpurun_out/h_c2_$name.err
  echo "$name rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/h_c2_$name.json').read().strip().splitlines()[-1])
    s=d['decode_stats_per_step']
    print('$name', round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'extra rounds', s['slow_path_symbols'], 'newton', s['consumer_polls'], 'bpp', round(d['bpp'],4), 'cnn TF', round(d['cnn_tflops'],1))
except Exception as e: print('$name failed', e)
PY
done
LLICTI_PROF_DUMP=1 timeout 300 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config > gpurun_out/h_c2_dump.json 2> gpurun_out/h_c2_dump.err
grep -E "class (6|8|1) " gpurun_out/h_c2_dump.err | tail -60 | awk '{printf "%s:%s ", $6, $7} END {print ""}'
NCU="ncu --set full --clock-control none --import-source on"
C2="python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu --no-per-config"
$NCU --kernel-name-base demangled -k "regex:cnn_tc_kernel<\(int\)2" -s 4 -c 1 -o gpurun_out/r02_cnn_c2_tmemH $C2 > gpurun_out/ncu_cnn_h.log 2>&1; echo "ncu cnn rc=$?"
C2F="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config"
$NCU -k regex:decode_band_lane -s 8 -c 1 -o gpurun_out/r02_lane_s0 $C2F > gpurun_out/ncu_lane_h.log 2>&1; echo "ncu lane rc=$?"
ls -la gpurun_out/*.ncu-rep
