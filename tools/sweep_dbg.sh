mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/dbg_$tag.json 2> gpurun_out/dbg_$tag.err; }
run c4 LLICTI_WAVE_CHAINS_PER_CTA=4
run c8 LLICTI_WAVE_CHAINS_PER_CTA=8
run c4s LLICTI_WAVE_CHAINS_PER_CTA=4 LLICTI_WAVE_SHARE_SMS=1
run c8s LLICTI_WAVE_CHAINS_PER_CTA=8 LLICTI_WAVE_SHARE_SMS=1
python - <<'PY'
import json
for c in ["c4","c8","c4s","c8s"]:
    try:
        d=json.load(open("gpurun_out/dbg_%s.json"%c)); s=d["decode_stats_per_step"]
        items=int(s["slow_path_symbols"])>>32
        print(c, round(d["kernel_ms_per_step"]["decode"],1), "cons cyc/run %.2fM wait %.0f%% redo %.0f%%"%(s["consumer_cycles"]/s["consumer_runs"]/1e6, 100*s["consumer_wait_cycles"]/s["consumer_cycles"], 100*(s["consumer_redo_cycles"]%1e10)/s["consumer_cycles"]), "prod busy/item %.0f cyc, busy frac %.2f"%(s["stat2"]/items, s["stat2"]/s["consumer_redo_cycles"]))
    except Exception as e: print(c, "ERR", e)
PY
