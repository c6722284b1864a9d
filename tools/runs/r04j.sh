mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "vit_attention_tcgen05" 2>&1 | tail -5
XFM_ATTN_PROF=1 timeout 300 python tools/attn_case.py fwd 1 2>&1 | tail -3
timeout 300 python tools/dev_kernels.py attn 2>&1 | grep -E "vit_self_tcgen05|vqkd_self_tcgen05" | tee gpurun_out/r04j_dev_attn.jsonl
