mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r04t_pytest.log; tail -4 gpurun_out/r04t_pytest.log
timeout 900 python bench.py > gpurun_out/r04t_bench.json 2> gpurun_out/r04t_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r04t_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r04t_bench.json")); print("pretrain", d["value"], d["ms_per_step"], d["e2e"], d["launch_sequence"]["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["frac"], d["fusion_layer"], d.get("gpu_eager_baseline"), d.get("cpu_baseline"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r04t_bench_reference.json 2> gpurun_out/r04t_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r04t_bench_reference.json | cut -c1-400
