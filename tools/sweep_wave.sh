mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "round_trip or piped or full_size or golden or rate" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest.log
for cfg in 5 4 3; do export LLICTI_WAVE_PRODUCER_CTAS_PER_SM=$cfg; timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/w_$cfg.json 2> gpurun_out/w_$cfg.err; done
python - <<'PY'
import json
for c in ["5","4","3"]:
    try:
        d=json.load(open("gpurun_out/w_%s.json"%c)); print(c, round(d["value"],1), round(d["decode_ms_per_step"],1), round(d["kernel_ms_per_step"]["decode"],1), d["decode_stats_per_step"]["consumer_polls"])
    except Exception as e: print(c, "ERR", e)
PY
tail -n 2 gpurun_out/w_*.err
