import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_model_gpu as T
from xfm_b200 import lib as L
g = T._load('/root/repo/tests/golden', 'base_mse.pt')
for mode in ('tc', 'no_cross_tc', 'no_tc_at_all'):
    orig_f, orig_b = L.attention_fwd, L.attention_bwd
    if mode != 'tc':
        def f(*a, _o=orig_f, **k):
            if mode == 'no_tc_at_all' or k.get('kv_samples') is not None: k['allow_tc'] = False
            return _o(*a, **k)
        L.attention_fwd = f
    model, cfg = T._build(g)
    batch = T._batch(g, cfg)
    with torch.no_grad():
        out = T._run(model, g, batch)
    print(mode, {k: (round(float(out[k]), 6), round(v, 6), f"{abs(float(out[k]) - v) / max(1, abs(v)):.2e}") for k, v in g['losses'].items()})
    L.attention_fwd = orig_f
