#!/bin/bash
# What the driver runs at round end: the full GPU suite, smoke(), the default bench line, the reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'enc', round(d['encode_mpps']), 'dec', round(d['decode_mpps']), {k:round(v,2) for k,v in d['kernel_ms_per_step'].items()})
    print('roofline', {k:(round(v,4) if isinstance(v,float) else v) for k,v in d['roofline'].items() if k in ('kernel','bound','achieved','peak','frac','traffic','share_of_step')})
    print('cpu', d.get('cpu_baseline')); print('per_config', {k:(round(v['value'],1) if 'value' in v else v) for k,v in d.get('per_config',{}).items()})
    print('train', d['per_config'].get('train_step'))
except Exception as e: print('failed', e)
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_reference_arm.err; echo "reference arm rc=$?"; tail -c 400 gpurun_out/bench_reference_arm.json
