timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python tools/dev_kernels.py attn 2>&1 | grep -E "vit_self_tcgen05|vqkd_self_tcgen05"
timeout 300 python tools/diag_balance.py 2>&1 | tail -4
