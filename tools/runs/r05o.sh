mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
for c in retrieval nlvr vqa; do
timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-eager --no-cpu > gpurun_out/r05o_bench_$c.json 2> gpurun_out/r05o_bench_$c.err; echo "$c rc=$?"
done
python - <<PY
import json
for c in ("retrieval","nlvr","vqa"):
    d=json.load(open(f"gpurun_out/r05o_bench_{c}.json")); print(c, d["value"], d["unit"], d["ms_per_step"])
PY
