#!/bin/bash
# Round 2, call f: vectorised colour kernels (parity), TMEM read-out probe, full ncu captures of the CNN (band 2, scale 0) and of the
# group decoder at the full c2 batch.
mkdir -p gpurun_out
tools/_bin/tmem_probe > gpurun_out/tmem_probe.txt 2>&1; echo "tmem rc=$?"; cat gpurun_out/tmem_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_f.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_f.log
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config > gpurun_out/f_c2.json 2> gpurun_out/f_c2.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/f_c2.json').read().strip().splitlines()[-1])
print(round(d['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
PY
NCU="ncu --set full --clock-control none --import-source on"
C2="python bench.py --workload c2 --images 8 --steps 1 --warmup 1 --no-cpu --no-per-config"
$C2 > gpurun_out/plain_f.log 2>&1 && $NCU --kernel-name-base demangled -k "regex:cnn_tc_kernel<\(int\)2" -s 0 -c 5 -o gpurun_out/r02_cnn_c2 $C2 > gpurun_out/ncu_cnn_f.log 2>&1; echo "ncu cnn rc=$?"
C2F="python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu --no-per-config"
$NCU -k regex:decode_band_group -s 0 -c 15 -o gpurun_out/r02_group_full $C2F > gpurun_out/ncu_group_f.log 2>&1; echo "ncu group rc=$?"
