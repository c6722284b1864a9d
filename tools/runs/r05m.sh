mkdir -p gpurun_out
timeout 400 python tools/bench_retrieval_eval.py --res 224 2>&1 | tail -1
timeout 400 python tools/bench_retrieval_eval.py --res 384 2>&1 | tail -1
cp gpurun_out/retrieval_eval.jsonl gpurun_out/r05m_retrieval_eval.jsonl
