mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/sh_$tag.json 2> gpurun_out/sh_$tag.err; }
run base LLICTI_WAVE_SHARE_SMS=0
run share LLICTI_WAVE_SHARE_SMS=1
run share_p4 LLICTI_WAVE_SHARE_SMS=1 LLICTI_WAVE_PRODUCER_CTAS_PER_SM=4
run share_p8 LLICTI_WAVE_SHARE_SMS=1 LLICTI_WAVE_PRODUCER_CTAS_PER_SM=8
run base_p8 LLICTI_WAVE_SHARE_SMS=0 LLICTI_WAVE_PRODUCER_CTAS_PER_SM=8
python - <<'PY'
import json
for c in ["base","share","share_p4","share_p8","base_p8"]:
    try:
        d=json.load(open("gpurun_out/sh_%s.json"%c)); print(c, round(d["value"],1), round(d["decode_ms_per_step"],1), round(d["kernel_ms_per_step"]["decode"],1), d["decode_stats_per_step"]["consumer_polls"])
    except Exception as e: print(c, "ERR", e)
PY
tail -n 2 gpurun_out/sh_*.err
