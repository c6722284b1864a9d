mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_model_gpu.py -m gpu -q -x -k "vq" 2>&1 | tail -5
python - <<PY
import torch, sys
sys.path.insert(0, ".")
from xfm_b200 import lib as L
z = torch.randn(96*196, 32, device="cuda"); cb = torch.nn.functional.normalize(torch.randn(8192, 32, device="cuda"), dim=-1)
for _ in range(3): L.vq_argmin(z, cb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): L.vq_argmin(z, cb)
e1.record(); torch.cuda.synchronize()
print("vq_argmin us", e0.elapsed_time(e1) * 1000 / 20)
PY
