mkdir -p gpurun_out
for c in 4 6 8 10; do export LLICTI_WAVE_CHAINS_PER_CTA=$c; timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/cpc_$c.json 2> gpurun_out/cpc_$c.err; done
python - <<'PY'
import json
for c in [4,6,8,10]:
    try:
        d=json.load(open("gpurun_out/cpc_%d.json"%c)); print(c, round(d["value"],1), round(d["decode_ms_per_step"],1), round(d["kernel_ms_per_step"]["decode"],1), round(d["kernel_ms_per_step"]["cnn"],1), d["decode_stats_per_step"]["consumer_polls"], d["gpu_launches"])
    except Exception as e: print(c, "ERR", e)
PY
tail -n 2 gpurun_out/cpc_*.err
