mkdir -p gpurun_out
timeout 300 python tools/dev_gemm_f32epi.py zsmem > gpurun_out/r02o_f32epi.log 2>&1; echo "rc=$?"; tail -9 gpurun_out/r02o_f32epi.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02o_pytest.log; tail -4 gpurun_out/r02o_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("r02o_bench",):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"])
PY
