mkdir -p gpurun_out
timeout 120 ./build/tmem_bench > gpurun_out/r04k_tmem_bench.jsonl 2>&1; echo "tmem rc=$?"; cat gpurun_out/r04k_tmem_bench.jsonl
