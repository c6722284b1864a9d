"""World-size-2 `gloo` tests (CPU) of the host logic of the data-parallel path:

  * XFMBase._gather_world == the reference's AllGather (models/xfm.py:81-101): rank-ordered features, local offset;
    the loss every rank computes on the gathered features equals the single-process big-batch loss, and the local
    gradient slice equals the big-batch gradient rows (the reference keeps only that slice, xfm.py:93-98);
  * B200DDPAccelerator on a real tiny model: flat broadcast at set_up; bucketed SUM all-reduce of the flat gradient buffer
    (averaging is the optimizer kernel's grad_mul = 1/W, i.e. DDP's mean) — plain, overlapped range by range, and with a
    second backward pass arriving after an early reduction (exact through the stash); "auto" overlap; the live-parameter
    table agreed over ranks.

The arithmetic checker is the oracle (test infrastructure); the CUDA kernels are not involved here.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        from oracle import xfm_oracle as O
        from xfm_b200.accelerator import B200DDPAccelerator
        from xfm_b200.xfm import XFMBase

        B, E = 6, 16
        g = torch.Generator().manual_seed(5)
        fi_all = torch.nn.functional.normalize(torch.randn(world * B, E, generator=g), dim=-1)
        ft_all = torch.nn.functional.normalize(torch.randn(world * B, E, generator=g), dim=-1)
        idx_all = torch.randint(0, 5, (world * B,), generator=g)
        temp = torch.tensor(0.07)
        res = {}
        for use_idx in (False, True):
            # ---- single-process big batch (what W ranks together must reproduce)
            a, b = fi_all.clone().requires_grad_(True), ft_all.clone().requires_grad_(True)
            big = O.contrastive_loss(a, b, temp, idx_all if use_idx else None)
            big.backward()
            # ---- this rank: gather through the product's host code, loss on the gathered features, local slice grad
            sl = slice(rank * B, (rank + 1) * B)
            fi, ft = fi_all[sl].clone(), ft_all[sl].clone()
            ia, ta, ix, off = XFMBase._gather_world(None, fi, ft, idx_all[sl].clone() if use_idx else None)
            assert off == rank * B
            assert torch.equal(ia, fi_all) and torch.equal(ta, ft_all)
            assert (ix is None) == (not use_idx) and (ix is None or torch.equal(ix, idx_all))
            ia, ta = ia.clone().requires_grad_(True), ta.clone().requires_grad_(True)
            loss = O.contrastive_loss(ia, ta, temp, ix)
            loss.backward()
            res[use_idx] = (float(loss) - float(big), float((ia.grad[sl] - a.grad[sl]).abs().max()),
                            float((ta.grad[sl] - b.grad[sl]).abs().max()))

        # ---- gradient exchange of the accelerator on a real (tiny) model: plain, overlapped per range, and with a second
        # backward pass arriving after ranges were reduced early (Pretrain.py:218-243 accumulates several per step)
        from xfm_b200.accelerator import FlatAdamW
        from xfm_b200.model_pretrain import XFM
        cfg = O.tiny_config(use_vision_tokenizer=True)
        model = XFM(dict(cfg), init=lambda n, s: O.make_tensor(n, s, rank), device="cpu")   # rank-dependent init ...
        opt = FlatAdamW(model, lr=1e-4)
        acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0, ALLREDUCE_BUCKETS=3, OVERLAP_ALLREDUCE=True))
        ref_P = XFM(dict(cfg), init=lambda n, s: O.make_tensor(n, s, 0), device="cpu").flat.P
        acc.set_up(model, opt, None, 0, world, rank)
        res["broadcast"] = float((model.flat.P - ref_P).abs().max())                           # ... equal after set_up
        G, n = model.flat.G, model.flat.G.numel()
        te = acc._train_end
        assert 0 < te < n and len(acc._blocks) == cfg["vision_depth"] and acc._vis[0] < acc._blocks[0][0]
        base = torch.arange(n, dtype=torch.float32) % 1000
        tot = sum(r + 1 for r in range(world))

        def frozen_ok():
            return float((G[te:] - base[te:] * (rank + 1)).abs().max())

        # (A) no overlap
        G.copy_(base * (rank + 1))
        acc._on_backward_begin()
        acc.all_reduce_grads(model)
        res["plain"] = float((G[:te] - base[:te] * tot).abs().max()) + frozen_ok()
        # (B) overlapped: non-vision ranges at the last node, vision blocks as they finish, tail at optimizer_step
        G.copy_(base * (rank + 1))
        acc._on_backward_begin()
        cb = acc._on_last_node()
        for i in reversed(range(len(acc._blocks))):
            cb(i)
        v0 = acc._vis[0]
        b0 = acc._blocks[0][0]
        res["early_done"] = float((G[:v0] - base[:v0] * tot).abs().max())                      # already summed
        res["tail_pending"] = float((G[v0:b0] - base[v0:b0] * (rank + 1)).abs().max())         # cls / patch embedding: not yet
        acc.all_reduce_grads(model)
        res["overlap"] = float((G[:te] - base[:te] * tot).abs().max()) + frozen_ok()
        # (C) a second backward pass after the early reduction: reduced values are stashed, the result stays exact
        G.copy_(base * (rank + 1))
        acc._on_backward_begin()
        cb = acc._on_last_node()
        for i in reversed(range(len(acc._blocks))):
            cb(i)
        acc._on_backward_begin()                      # backward #2 begins
        G[:te] += 0.5 * base[:te] * (rank + 2)        # ... and accumulates local gradients
        cb = acc._on_last_node()
        cb(len(acc._blocks) - 1)                      # only one block reduced early this time
        acc.all_reduce_grads(model)
        want = base[:te] * tot + 0.5 * base[:te] * sum(r + 2 for r in range(world))
        res["stash"] = float((G[:te] - want).abs().max()) + frozen_ok()
        # (D) "auto": learns the number of backward passes per optimizer step; the first step is never overlapped
        auto = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
        auto.set_up(model, opt, None, 0, world, rank)
        seen = []
        for step in range(3):
            for k in range(2):
                auto._on_backward_begin()
                seen.append(auto._on_last_node() is not None)
            auto.all_reduce_grads(model)
            auto._bw_hist = (auto._bw_hist + [auto._bw_seen])[-4:]
            auto._bw_per_step, auto._bw_seen = max(auto._bw_hist), 0
        res["auto"] = seen
        # (E) the live-parameter table is agreed over ranks (a parameter is updated if ANY rank has a gradient for it)
        t = torch.tensor([0, 255, 3, 255] if rank == 0 else [255, 255, 3, 1], dtype=torch.uint8)
        res["live"] = acc._sync_live(t).tolist()
        q.put((rank, res, None))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, None, traceback.format_exc()))


@pytest.mark.timeout(180)
def test_two_rank_gather_and_allreduce():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, res, err in out:
        assert err is None, f"rank {rank}:\n{err}"
        for use_idx in (False, True):
            dl, di, dt = res[use_idx]
            assert abs(dl) < 1e-6 and di < 1e-6 and dt < 1e-6, (rank, use_idx, res[use_idx])
        for k in ("broadcast", "plain", "early_done", "tail_pending", "overlap", "stash"):
            assert res[k] == 0.0, (rank, k, res[k])
        assert res["auto"] == [False, False, False, True, False, True], res["auto"]
        assert res["live"] == [0, 255, 3, 1]


def test_auto_overlap_policy_by_world_size():
    """"auto" overlaps the all-reduce with the backward pass on 2 ranks only (accelerator docstring: measured on 4 and 8 GPUs
    the exposed single all-reduce is faster); an explicit `true` overlaps at any size."""
    from xfm_b200.accelerator import B200DDPAccelerator
    for world, want in ((2, True), (4, False), (8, False)):
        acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0))
        acc.world, acc._bw_per_step, acc._bw_seen = world, 1, 1
        assert acc._last_backward_expected() is want
        forced = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0, OVERLAP_ALLREDUCE=True))
        forced.world = world
        assert forced._last_backward_expected() is True
