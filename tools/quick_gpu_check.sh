mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest.log
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/b1.json 2> gpurun_out/b1.err
timeout 300 python bench.py --workload c2 --images 8 --steps 2 --warmup 1 --no-cpu > gpurun_out/b2.json 2> gpurun_out/b2.err
python - <<'PY'
import json
for f in ["gpurun_out/b1.json","gpurun_out/b2.json"]:
    d=json.load(open(f)); print(round(d["value"],1), round(d["encode_ms_per_step"],2), round(d["decode_ms_per_step"],2), {k:round(v,2) for k,v in d["kernel_ms_per_step"].items()}, round(d["cnn_tflops"],1), d["decode_stats_per_step"])
PY
tail -n 3 gpurun_out/b1.err gpurun_out/b2.err
