mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "vit_attention_tcgen05_forward" 2>&1 | tail -8
for sp in 1 0; do XFM_VIT_FWD_SP=$sp timeout 300 python tools/dev_kernels.py attn 2>&1 | grep -E "vit_self_tcgen05|vqkd_self_tcgen05" | sed "s/^/sp=$sp /"; done
