mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "vit_attention_tcgen05" 2>&1 | tail -15
