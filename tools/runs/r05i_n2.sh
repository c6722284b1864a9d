mkdir -p gpurun_out
export XFM_BENCH_WATCHDOG=150
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -3
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-eager --no-cpu > gpurun_out/r05i_n2.json 2> gpurun_out/r05i_n2.err; echo "n2 rc=$? lines=$(wc -l < gpurun_out/r05i_n2.json)"; tail -2 gpurun_out/r05i_n2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r05i_n2.json")); print("n2", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("launch_sequence", {}).get("ms_per_step"))
PY
