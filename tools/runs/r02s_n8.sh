mkdir -p gpurun_out
export XFM_BENCH_WATCHDOG=150
run() {  # tag nproc extra...
  tag=$1; n=$2; shift 2
  timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/r02s_$tag.json 2> gpurun_out/r02s_$tag.err
  echo "$tag rc=$? lines=$(wc -l < gpurun_out/r02s_$tag.json)"; tail -c 300 gpurun_out/r02s_$tag.err | tail -2
}
run n8_plain 8 --overlap false
run n8_true 8 --overlap true
run n4_plain 4 --overlap false
NCCL_ALGO=NVLS run n8_plain_nvls 8 --overlap false
python - <<PY
import json
for f in ("n8_plain", "n8_true", "n4_plain", "n8_plain_nvls"):
    try:
        d=json.load(open(f"gpurun_out/r02s_{f}.json")); print(f, d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("launch_sequence", {}).get("ms_per_step"))
    except Exception as e:
        print(f, "unreadable", e)
PY
