#!/bin/bash
# Round 2, call i: CNN with implicit im2col (shifted A descriptors): full GPU suite, bench, ncu of the CNN.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cnn" > gpurun_out/pytest_i0.log 2>&1; echo "cnn pytest rc=$?"; tail -15 gpurun_out/pytest_i0.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_i.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_i.log
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/i_c2.json 2> gpurun_out/i_c2.err
echo "bench rc=$?"; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/i_c2.json').read().strip().splitlines()[-1])
    s=d['decode_stats_per_step']
    print(round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()}, 'bpp', round(d['bpp'],4), 'cnn TF', round(d['cnn_tflops'],1))
except Exception as e: print('failed', e)
PY
LLICTI_PROF_DUMP=1 timeout 300 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config > gpurun_out/i_c2_dump.json 2> gpurun_out/i_c2_dump.err
grep -E "class (6|8|1) " gpurun_out/i_c2_dump.err | tail -60 | awk '{printf "%s:%s ", $6, $7} END {print ""}'
NCU="ncu --set full --clock-control none --import-source on"
C2="python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu --no-per-config"
$NCU --kernel-name-base demangled -k "regex:cnn_tc_kernel<\(int\)2" -s 4 -c 1 -o gpurun_out/r02_cnn_c2_implicit $C2 > gpurun_out/ncu_cnn_i.log 2>&1; echo "ncu cnn rc=$?"
ls -la gpurun_out/*.ncu-rep
