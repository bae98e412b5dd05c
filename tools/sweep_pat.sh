mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/pat_$tag.json 2> gpurun_out/pat_$tag.err; }
for p in 122 111 211 322 221 0; do run $p LLICTI_WAVE_PATTERN=$p; done
run 111s LLICTI_WAVE_PATTERN=111 LLICTI_WAVE_SHARE_SMS=1
run 211s LLICTI_WAVE_PATTERN=211 LLICTI_WAVE_SHARE_SMS=1
python - <<'PY'
import json
for c in ["122","111","211","322","221","0","111s","211s"]:
    try:
        d=json.load(open("gpurun_out/pat_%s.json"%c)); s=d["decode_stats_per_step"]
        print(c, round(d["value"],1), round(d["kernel_ms_per_step"]["decode"],1), "cons cyc/run %.2fM wait %.0f%%"%(s["consumer_cycles"]/s["consumer_runs"]/1e6, 100*s["consumer_wait_cycles"]/s["consumer_cycles"]))
    except Exception as e: print(c, "ERR", e)
PY
