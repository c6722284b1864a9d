mkdir -p gpurun_out
for c in retrieval nlvr vqa; do
timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-eager --no-cpu > gpurun_out/r04o_bench_$c.json 2> gpurun_out/r04o_bench_$c.err; echo "$c rc=$?"; tail -2 gpurun_out/r04o_bench_$c.err
done
python - <<PY
import json
for c in ("retrieval","nlvr","vqa"):
    d=json.load(open(f"gpurun_out/r04o_bench_{c}.json")); print(c, d["value"], d["unit"], d["ms_per_step"], d.get("launch_sequence",{}).get("ms_per_step"), d.get("roofline",{}).get("step_tflops_algorithmic"))
PY
