mkdir -p gpurun_out
export XFM_BENCH_WATCHDOG=150
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -4
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02y_n2.json 2> gpurun_out/r02y_n2.err; echo "n2 rc=$? lines=$(wc -l < gpurun_out/r02y_n2.json)"
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --overlap false > gpurun_out/r02y_n2_plain.json 2> gpurun_out/r02y_n2_plain.err; echo "n2 plain rc=$?"
python - <<PY
import json
for f in ("n2", "n2_plain"):
    d=json.load(open(f"gpurun_out/r02y_{f}.json")); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("launch_sequence", {}).get("ms_per_step"))
PY
