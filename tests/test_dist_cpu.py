"""World-size-2 `gloo` tests (CPU) of the host logic of the data-parallel path:

  * XFMBase._gather_world == the reference's AllGather (models/xfm.py:81-101): rank-ordered features, local offset;
    the loss every rank computes on the gathered features equals the single-process big-batch loss, and the local
    gradient slice equals the big-batch gradient rows (the reference keeps only that slice, xfm.py:93-98);
  * B200DDPAccelerator.all_reduce_grads: bucketed SUM all-reduce of the flat gradient buffer (averaging is the optimizer
    kernel's grad_mul = 1/W, i.e. DDP's mean).

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

        # ---- gradient all-reduce of the trainable part of a flat buffer in buckets (frozen tail untouched)
        from collections import OrderedDict

        from xfm_b200.params import Segment

        class Flat:
            pass

        class Model:
            pass

        m = Model()
        m.flat = Flat()
        n = 64 * 37 + 64
        segs = OrderedDict()
        segs["temp"] = Segment("temp", (), 0, 1, True)
        segs["vision_encoder.a"] = Segment("vision_encoder.a", (64 * 10,), 64, 64 * 10, True)
        segs["text_encoder.b"] = Segment("text_encoder.b", (64 * 25 + 3,), 64 * 11, 64 * 25 + 3, True)
        segs["vqkd.frozen"] = Segment("vqkd.frozen", (64,), 64 * 37, 64, False)
        m.flat.segments = segs
        m.flat.G = torch.arange(n, dtype=torch.float32) * (rank + 1)
        acc = B200DDPAccelerator(dict(CLIP_GRAD_NORM=1.0, ALLREDUCE_BUCKETS=4))
        acc.world, acc.rank = world, rank
        acc._layout(m)
        assert acc._train_end == 64 * 37 and acc._vis == (64, 64 * 11)
        acc.all_reduce_grads(m)
        want = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
        want[64 * 37:] = torch.arange(64 * 37, n, dtype=torch.float32) * (rank + 1)   # frozen segment: not reduced
        res["allreduce"] = float((m.flat.G - want).abs().max())
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
        assert res["allreduce"] == 0.0
