mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/kn_$tag.json 2> gpurun_out/kn_$tag.err; }
run s0 LLICTI_WAVE_SHARE_SMS=0
run s4 LLICTI_WAVE_SHARE_SMS=4
run s8 LLICTI_WAVE_SHARE_SMS=8
run s12 LLICTI_WAVE_SHARE_SMS=12
run s99 LLICTI_WAVE_SHARE_SMS=99
python - <<'PY'
import json
for c in ["s0","s4","s8","s12","s99"]:
    try:
        d=json.load(open("gpurun_out/kn_%s.json"%c)); s=d["decode_stats_per_step"]
        print(c, round(d["value"],1), round(d["decode_ms_per_step"],2), "wait %.0f%%"%(100*s["consumer_wait_cycles"]/s["consumer_cycles"]), "cyc/run %.2fM"%(s["consumer_cycles"]/s["consumer_runs"]/1e6))
    except Exception as e: print(c, "ERR", e)
PY
