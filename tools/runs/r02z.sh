mkdir -p gpurun_out
timeout 300 python tools/dev_kernels.py ln > gpurun_out/r02z_dev_ln.log 2>&1; tail -4 gpurun_out/r02z_dev_ln.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02z_pytest.log; tail -3 gpurun_out/r02z_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r02z_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["fusion_layer"]["ms_per_step"])
PY
