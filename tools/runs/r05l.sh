timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "vit_attention" 2>&1 | tail -3
timeout 300 python tools/dev_kernels.py attn 2>&1 | grep -E "vit_self_tcgen05|vqkd_self_tcgen05"
timeout 300 python tools/time_vit577.py 32
