mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02u_pytest.log; tail -4 gpurun_out/r02u_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02u_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02u_smoke.log
timeout 900 python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02u_bench_reference.json 2> gpurun_out/r02u_bench_reference.err; echo "ref rc=$?"; head -c 600 gpurun_out/r02u_bench_reference.json
for c in retrieval nlvr vqa; do timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r02u_bench_$c.json 2> gpurun_out/r02u_bench_$c.err; echo "$c rc=$?"; done
python - <<PY
import json
d=json.load(open("gpurun_out/r02u_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d.get("gpu_eager_baseline",{}).get("bf16_autocast"), d["cpu_baseline"])
for c in ("retrieval","nlvr","vqa"):
    d=json.load(open(f"gpurun_out/r02u_bench_{c}.json")); print(c, d["value"], d["ms_per_step"], d.get("launch_sequence",{}).get("ms_per_step"))
PY
