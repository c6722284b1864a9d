mkdir -p gpurun_out
timeout 120 ./build/tmem_bench > gpurun_out/r04a_tmem_bench.jsonl 2>&1; echo "tmem rc=$?"; cat gpurun_out/r04a_tmem_bench.jsonl
XFM_ATTN_PROF=1 timeout 300 python tools/attn_case.py fwd 2 > gpurun_out/r04a_attn_prof.txt 2>&1; echo "prof rc=$?"; tail -8 gpurun_out/r04a_attn_prof.txt
