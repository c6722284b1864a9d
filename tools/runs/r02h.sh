mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5"
timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q 2>&1 | tail -5
timeout 300 $T > gpurun_out/r02h_n2.json 2> gpurun_out/r02h_n2.err; tail -2 gpurun_out/r02h_n2.err
timeout 400 $T --graph > gpurun_out/r02h_n2_graph.json 2> gpurun_out/r02h_n2_graph.err; echo "graph rc=$?"; tail -5 gpurun_out/r02h_n2_graph.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-eager --no-cpu --graph > gpurun_out/r02h_n1.json 2> gpurun_out/r02h_n1.err
python - <<PY
import json
for f in ("n1","n2","n2_graph"):
    try:
        d=json.loads(open("gpurun_out/r02h_%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],2), d["step_ms"], d.get("cuda_graph"))
    except Exception as e: print(f, "ERR", e)
PY
