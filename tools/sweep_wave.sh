mkdir -p gpurun_out
for cfg in 32_8 16_16 16_8 64_4 8_32; do export LLICTI_WAVE_STRIP_ROWS=${cfg%_*} LLICTI_WAVE_MAX_STRIPS=${cfg#*_}; timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/w_$cfg.json 2> gpurun_out/w_$cfg.err; done
python - <<'PY'
import json
for c in ["32_8","16_16","16_8","64_4","8_32"]:
    try:
        d=json.load(open("gpurun_out/w_%s.json"%c)); print(c, round(d["value"],1), round(d["decode_ms_per_step"],1), round(d["kernel_ms_per_step"]["decode"],1), round(d["kernel_ms_per_step"]["cnn"],1), d["decode_stats_per_step"]["consumer_polls"], d["gpu_launches"])
    except Exception as e: print(c, "ERR", e)
PY
tail -n 2 gpurun_out/w_*.err
