mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/kn_$tag.json 2> gpurun_out/kn_$tag.err; }
run base LLICTI_X=0
run yahead LLICTI_WAVE_Y_AHEAD=1
run share LLICTI_WAVE_SHARE_SMS=1
run yahead_share LLICTI_WAVE_Y_AHEAD=1 LLICTI_WAVE_SHARE_SMS=1
run p211 LLICTI_WAVE_PATTERN=211
python - <<'PY'
import json
for c in ["base","yahead","share","yahead_share","p211"]:
    try:
        d=json.load(open("gpurun_out/kn_%s.json"%c)); s=d["decode_stats_per_step"]
        print(c, round(d["value"],1), round(d["decode_ms_per_step"],2), "wait %.0f%%"%(100*s["consumer_wait_cycles"]/s["consumer_cycles"]))
    except Exception as e: print(c, "ERR", e)
PY
