#!/bin/bash
# Round-2 evidence run: full GPU suite, the default bench line (c2 + per_config + CPU reference leg), the reference arm, launch list
# and one full ncu capture per major kernel of the same command.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
    print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('kernel','bound','achieved','peak','frac','traffic','share_of_step')})
    print('cpu', d.get('cpu_baseline')); print('per_config', {k:round(v['value'],1) for k,v in d.get('per_config',{}).items()})
except Exception as e: print('failed', e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_reference_arm.json
timeout 300 python bench.py --batch 50 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/bench_c2_batch50.json 2> gpurun_out/bench_c2_batch50.err; echo "batch50 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c2_batch50.json').read().strip().splitlines()[-1])
    print('batch50 value', round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
except Exception as e: print('failed', e)
PY
C2="python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu --no-per-config"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches_c2_8img.csv $C2 > gpurun_out/ncu_launches.log 2>&1; echo "launch list rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
$NCU --kernel-name-base demangled -k "regex:cnn_tc_kernel<\(int\)2" -s 4 -c 1 -o gpurun_out/${TAG}_cnn_c2 $C2 > gpurun_out/ncu_a.log 2>&1; echo "ncu cnn rc=$?"
$NCU -k regex:decode_band_lane -s 3 -c 1 -o gpurun_out/${TAG}_lane_c2 $C2 > gpurun_out/ncu_b.log 2>&1; echo "ncu lane rc=$?"
$NCU -k regex:band_bounds -s 12 -c 1 -o gpurun_out/${TAG}_bounds_c2 $C2 > gpurun_out/ncu_c.log 2>&1; echo "ncu bounds rc=$?"
ls -la gpurun_out/${TAG}_*
