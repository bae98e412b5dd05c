# Full bench lines of every BASELINE.json configuration (one GPU): gpurun_out/full_c*.json
mkdir -p gpurun_out
( time python bench.py > gpurun_out/full_c1.json 2> gpurun_out/full_c1.err ) 2> gpurun_out/full_c1.time
for w in c0 c2 c3 c4; do timeout 600 python bench.py --workload $w --no-cpu > gpurun_out/full_$w.json 2> gpurun_out/full_$w.err; done
( time python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/full_ref.json 2> gpurun_out/full_ref.err ) 2> gpurun_out/full_ref.time
python - <<'PY'
import json
for w in ["c1","c0","c2","c3","c4","ref"]:
    try:
        d=json.loads(open("gpurun_out/full_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["value"],2), d.get("encode_mpps"), d.get("decode_mpps"), d.get("e2e",{}).get("value"), d.get("cpu_baseline"))
    except Exception as e: print(w, "ERR", e)
PY
cat gpurun_out/full_c1.time gpurun_out/full_ref.time; tail -n 2 gpurun_out/full_*.err
