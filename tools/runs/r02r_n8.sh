mkdir -p gpurun_out
export XFM_BENCH_WATCHDOG=150
nvidia-smi -L | head -8
run() {  # tag nproc extra...
  tag=$1; n=$2; shift 2
  timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/r02r_$tag.json 2> gpurun_out/r02r_$tag.err
  echo "$tag rc=$? lines=$(wc -l < gpurun_out/r02r_$tag.json)"; tail -c 600 gpurun_out/r02r_$tag.err | tail -3
}
run n8 8
run n8_nograph 8 --no-graph
run n4 4
python - <<PY
import json
for f in ("n8", "n8_nograph", "n4"):
    try:
        d=json.load(open(f"gpurun_out/r02r_{f}.json")); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("launch_sequence", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
