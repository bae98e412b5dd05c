mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 120 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/kn_$tag.json 2> gpurun_out/kn_$tag.err; }
for p in 122 100 311 522 133; do run p$p LLICTI_WAVE_PATTERN=$p; done
python - <<'PY'
import json
for c in ["p122","p100","p311","p522","p133"]:
    try:
        d=json.load(open("gpurun_out/kn_%s.json"%c)); s=d["decode_stats_per_step"]
        print(c, round(d["value"],1), round(d["decode_ms_per_step"],2), "wait %.0f%%"%(100*s["consumer_wait_cycles"]/s["consumer_cycles"]), "cyc/run %.2fM"%(s["consumer_cycles"]/s["consumer_runs"]/1e6))
    except Exception as e: print(c, "ERR", e)
PY
