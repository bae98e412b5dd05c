#!/bin/bash
# Lane decoder check: its tests, then the c2 bench line with the synthetic stand-in weights and with the trained ones.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane or round_trip or full_size" > gpurun_out/pytest_q.log 2>&1 || { tail -8 gpurun_out/pytest_q.log; echo "tests failed"; exit 1; }
tail -1 gpurun_out/pytest_q.log
for w in "" "--weights-npz tests/golden/ckpt_A_trained.npz"; do
timeout 300 python bench.py --workload c2 --steps 3 --warmup 2 --no-cpu --no-per-config $w > gpurun_out/ql_c2.json 2> gpurun_out/ql_c2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/ql_c2.json").read().strip().splitlines()[-1])
print("weights [$w]: value", round(d["value"],1), "dec", round(d["decode_mpps"]), "decode kernel", round(d["kernel_ms_per_step"]["decode"],2), "bpp", round(d["bpp"],3), "newton", d["decode_stats_per_step"]["consumer_polls"], "extra", d["decode_stats_per_step"]["slow_path_symbols"])
PY
done
