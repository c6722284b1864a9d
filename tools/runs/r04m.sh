mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" 2>&1 | tail -5
timeout 300 python tools/dev_kernels.py attn 2>&1 | grep -E "tcgen05|fusion_cross" | tee gpurun_out/r04m_dev_attn.jsonl
