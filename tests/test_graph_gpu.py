"""Whole-step CUDA-graph capture (xfm_b200/graph.py): a replayed step computes what the eager launch sequence computes, the
learning-rate schedule and fresh inputs reach the captured kernels, and dropout masks change from replay to replay through
the device-resident seed salt."""
import pytest
import torch

from oracle import xfm_oracle as O

pytestmark = pytest.mark.gpu


def _setup(train=False):
    from xfm_b200.accelerator import B200DDPAccelerator, FlatAdamW
    from xfm_b200.model_pretrain import XFM
    cfg = O.tiny_config()
    model = XFM(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cuda")
    model.train(train)
    opt = FlatAdamW(model, lr=1e-3, weight_decay=0.01, lr_mult=2.0)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0 / (1 + s))
    acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
    wrapped, opt, _ = acc.set_up(model, opt, None, 0, 1, 0)
    return cfg, model, wrapped, opt, sched, acc


def _loss_fn(mim):
    def fn(model, b):
        out = model(b["image"], b["text_ids"], b["text_atts"], text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"],
                    masked_ids=b["masked_ids"], ret_mim_loss=mim, data_source="image")
        return out["loss_itc"] + out["loss_itm"] + out["loss_mlm"] + (out["loss_mim"] if mim else 0.0)
    return fn


def test_graphed_step_equals_eager_step():
    from xfm_b200.graph import GraphedStep
    B = 4
    batches = [{k: v.cuda() for k, v in O.make_batch(O.tiny_config(), B, L=24, M=6, seed=s).items()} for s in (1, 2, 3)]
    neg = (torch.roll(torch.arange(B), 1).cuda(), torch.roll(torch.arange(B), -1).cuda())
    # eager: 3 warm-up steps on batch 0 (what GraphedStep runs before capturing), then batches 0, 1, 2
    cfg, model, wrapped, opt, sched, acc = _setup()
    model._forced_negatives = neg
    fn = _loss_fn(False)
    eager = []
    for i, b in enumerate([batches[0]] * 3 + batches):
        loss = fn(wrapped, b)
        acc.backward_step(loss, opt)
        acc.optimizer_step(opt, wrapped)
        if i >= 3:
            eager.append(float(loss.detach()))
            sched.step()
    p_eager = model.flat.P.clone()
    # graphed
    cfg, model, wrapped, opt, sched, acc = _setup()
    model._forced_negatives = neg
    step = GraphedStep(wrapped, opt, acc, fn, batches[0], warmup=3)   # 3 real warm-up steps (capturing executes nothing)
    graphed = []
    for b in batches:
        out = step(b)
        graphed.append(float(out[0]))
        sched.step()
    for a, g in zip(eager, graphed):
        assert abs(a - g) <= 2e-3 * abs(a), (eager, graphed)
    # parameters agree up to the order of fp32 atomic gradient accumulation (Adam turns noise-level gradients into +-lr)
    diff = (model.flat.P - p_eager).abs()
    assert float(diff.max()) <= 2.5e-2 and float((diff > 1e-4).float().mean()) < 0.05
    assert graphed[0] != graphed[1] and abs(eager[2] - eager[0]) > 1e-4      # the replays saw the new inputs


def test_graphed_train_step_with_mim_masks_and_fresh_dropout():
    from xfm_b200 import lib as L
    from xfm_b200.graph import GraphedStep
    B = 4
    b = {k: v.cuda() for k, v in O.make_batch(O.tiny_config(), B, L=24, M=6, seed=1).items()}
    cfg, model, wrapped, opt, sched, acc = _setup(train=True)
    for g in opt.param_groups:
        g["lr"] = 0.0                     # frozen weights: only masks / negatives / dropout differ between replays
    step = GraphedStep(wrapped, opt, acc, _loss_fn(True), b, uses_mim_masks=True)
    n0 = L.launch_count()
    losses = [float(step(b)[0]) for _ in range(4)]
    assert all(torch.isfinite(torch.tensor(losses))) and len(set(losses)) == 4, losses   # a fresh draw per replay
    assert L.launch_count() == n0          # nothing is launched from Python any more: one graph launch per step
    # the salt is what changes the masks: same seed, same index -> same mask until the device word advances
    x = torch.ones(1 << 14, device="cuda")
    L.seed_salt_set(0)
    m0 = L.dropout_apply(x, 0.5, 1234).float()
    m1 = L.dropout_apply(x, 0.5, 1234).float()
    L.seed_salt_bump(1)
    m2 = L.dropout_apply(x, 0.5, 1234).float()
    L.seed_salt_set(0)
    m3 = L.dropout_apply(x, 0.5, 1234).float()
    assert torch.equal(m0, m1) and torch.equal(m0, m3) and not torch.equal(m0, m2)
    assert abs(float(m2.mean()) - 1.0) < 0.05
