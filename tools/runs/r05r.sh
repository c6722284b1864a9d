timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_graph_gpu.py -m gpu -q 2>&1 | tail -3
