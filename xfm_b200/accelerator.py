"""Data-parallel accelerator for the B200 path — the interface of accelerators/accelerator.py:15-32 /
ddp_accelerator.py:34-98 (set_up / backward_step / optimizer_step, `.cfg.AUTO_CAST`), registered by a reference
checkout as ACCELERATOR_MAP['B200DDP'].

What changes underneath:
  * gradients of the whole model live in ONE flat fp32 buffer (params.FlatParams.G), so the DDP gradient exchange is a
    few large NCCL all-reduces (NVLS in-switch reduction on an NVSwitch box) instead of ~750 bucketed tensors with
    find_unused_parameters bookkeeping; it is issued once per optimizer step (the reference reduces after every
    backward, up to 5x per step with gradient accumulation across data streams, Pretrain.py:218-243);
  * clip_grad_norm_ + transformers.AdamW.step + zero_grad collapse into two kernels over the flat buffers
    (xfm_grad_sumsq, xfm_adamw_flat) that also refresh the bf16 weight shadow — no per-tensor launches, no host sync
    unless the caller asks for the norm as a float.
"""
import torch
import torch.distributed as dist

from . import lib as L

def reserve_arena(nbytes=None, factor=1.5, device=None):
    """Make the caching allocator own one large segment so steady-state steps never reach cudaMalloc.

    A pre-training step allocates ~2000 activation tensors with staggered lifetimes; the allocator keeps growing by a
    few hundred MB for dozens of steps, and every cudaMalloc that lands while the GPU is busy stalls the launching
    thread (measured on B200: 175 ms/step with 9 device allocations in a 6-step loop vs 111 ms/step with none).  Call
    after one warm-up step: reserves factor x the peak seen so far (or nbytes) in a single block and returns it to the
    cache, from which later allocations are carved."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    torch.cuda.synchronize(dev)
    if nbytes is None:
        nbytes = int(torch.cuda.max_memory_allocated(dev) * factor)
    free, _ = torch.cuda.mem_get_info(dev)
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info(dev)
    nbytes = min(nbytes, int(free * 0.9))
    if nbytes > (1 << 20):
        x = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        del x
    return nbytes


NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight", "norm.bias", "norm.weight", "norm1.bias", "norm1.weight",
            "norm2.bias", "norm2.weight")  # optim.py:17-25


class FlatAdamW:
    """optim.py:4-50 (4 parameter groups: {decay, no-decay} x {lr, lr * lr_mult}) over the flat parameter buffer.
    Exposes `param_groups` with an 'lr' entry per group so scheduler.py's LambdaLR-style schedulers can drive it."""

    def __init__(self, model, lr=1e-4, weight_decay=0.01, lr_mult=1.0, betas=(0.9, 0.98), eps=1e-8, correct_bias=True):
        self.model, self.flat = model, model.flat
        fp = self.flat
        self.betas, self.eps, self.correct_bias = betas, eps, correct_bias
        large = set(getattr(model, "init_params", []))
        self.param_groups = [dict(lr=lr, weight_decay=weight_decay, initial_lr=lr, params=[]),
                             dict(lr=lr, weight_decay=0.0, initial_lr=lr, params=[]),
                             dict(lr=lr * lr_mult, weight_decay=weight_decay, initial_lr=lr * lr_mult, params=[]),
                             dict(lr=lr * lr_mult, weight_decay=0.0, initial_lr=lr * lr_mult, params=[])]
        self.defaults = dict(lr=lr)
        self._group_of = {}
        for name, seg in fp.segments.items():
            if not seg.trainable or name.rsplit(".", 1)[-1].startswith("_"):
                continue
            gi = (1 if any(nd in name for nd in NO_DECAY) else 0) + (2 if name in large else 0)
            self._group_of[name] = gi
        n = fp.P.numel()
        self.M = torch.zeros(n, dtype=torch.float32, device=fp.P.device)
        self.V = torch.zeros(n, dtype=torch.float32, device=fp.P.device)
        self._chunks = {}
        self.step_count = 0
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=fp.P.device)
        self.norm = torch.zeros(1, dtype=torch.float32, device=fp.P.device)
        self.state = {}

    def _chunk_table(self):
        """uint8 group id per 64-element chunk; 255 for frozen segments and for parameters that received no gradient
        since the last zero_grad (the reference's optimizers skip grad-None parameters)."""
        key = frozenset(self.flat.touched)
        tab = self._chunks.get(key)
        if tab is None:
            fp = self.flat
            host = torch.full((fp.P.numel() // 64,), 255, dtype=torch.uint8)
            for name, gi in self._group_of.items():
                if name in fp.touched:
                    s = fp.segments[name]
                    host[s.offset // 64:(s.offset + s.numel + 63) // 64] = gi
            tab = host.to(fp.P.device)
            self._chunks = {key: tab}
        return tab

    def step(self, max_grad_norm=0.0, grad_mul=1.0):
        self.step_count += 1
        hp = L.AdamWParams()
        for i, g in enumerate(self.param_groups):
            hp.lr[i], hp.weight_decay[i] = g["lr"], g["weight_decay"]
        hp.beta1, hp.beta2, hp.eps = self.betas[0], self.betas[1], self.eps
        hp.max_grad_norm, hp.grad_mul, hp.step, hp.correct_bias = max_grad_norm, grad_mul, self.step_count, int(self.correct_bias)
        tab = self._chunk_table()
        fp = self.flat
        fresh = fp._shadow_version == fp.P._version
        L.grad_sumsq(fp.G, tab, self.sumsq)
        L.adamw_flat(fp.P, fp.G, self.M, self.V, fp.S, tab, hp, sumsq=self.sumsq, norm_out=self.norm)
        if fresh:
            fp._shadow_version = fp.P._version  # the kernel refreshed the bf16 shadow of everything it changed
        return self.norm

    def zero_grad(self, set_to_none=True):
        self.model.zero_grad()

    def state_dict(self):
        return dict(step=self.step_count, M=self.M, V=self.V, param_groups=[{k: v for k, v in g.items() if k != "params"}
                                                                            for g in self.param_groups])

    def load_state_dict(self, sd):
        self.step_count = sd["step"]
        self.M.copy_(sd["M"])
        self.V.copy_(sd["V"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)


class _Cfg:
    def __init__(self, d):
        self.__dict__.update(d)


class _Wrapped(torch.nn.Module):
    """What set_up returns as `model`: callable like the DDP wrapper, `.module` is the bare model (Pretrain.py:261-263)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


class B200DDPAccelerator:
    def __init__(self, cfg, logger=None):
        self.cfg = cfg if hasattr(cfg, "CLIP_GRAD_NORM") or not isinstance(cfg, dict) else _Cfg(cfg)
        if not hasattr(self.cfg, "AUTO_CAST"):
            self.cfg.AUTO_CAST = False
        self.clip = float(getattr(self.cfg, "CLIP_GRAD_NORM", 0.0) or 0.0)
        self.world, self.rank = 1, 0
        self.buckets = int(getattr(self.cfg, "ALLREDUCE_BUCKETS", 4))
        self.overlap = bool(getattr(self.cfg, "OVERLAP_ALLREDUCE", True))
        self._comm_stream = None
        self._early = None
        self._vis = None

    def set_up(self, model, optimizer, lr_scheduler, local_rank=0, world_size=1, rank=0):
        self.world, self.rank = world_size, rank
        self._layout(model)
        if world_size > 1:
            assert dist.is_initialized(), "init torch.distributed (nccl) before set_up"
            flat = model.flat
            dist.broadcast(flat.P, src=0)  # one flat broadcast replaces ~750 per-tensor ones (ddp_accelerator.py:69-74)
            flat.sync_shadow(force=True)
            if self.overlap and self._vis is not None and flat.G.is_cuda:
                self._model = model
                self._comm_stream = torch.cuda.Stream(device=flat.G.device)
                model._last_node_hook = self._early_reduce
        return _Wrapped(model), optimizer, lr_scheduler

    def _layout(self, model):
        """Ranges of the flat gradient buffer: [0, train_end) holds every trainable segment (the frozen VQ-KD tokenizer sits
        behind it and is never reduced); [v0, v1) is the vision encoder, whose gradients are the last ones produced."""
        segs = model.flat.segments
        al = lambda n: (n + 63) // 64 * 64
        self._train_end = max([s.offset + al(s.numel) for s in segs.values() if s.trainable] or [0])
        vis = [s for s in segs.values() if s.name.startswith("vision_encoder.")]
        self._vis = None
        if vis:
            v0, v1 = min(s.offset for s in vis), max(s.offset + al(s.numel) for s in vis)
            inside = [s for s in segs.values() if v0 <= s.offset < v1]
            if len(inside) == len(vis):   # contiguous
                self._vis = (v0, v1)

    def backward_step(self, loss, optimizer):
        if self._early is not None:
            raise RuntimeError("B200DDPAccelerator: a second backward ran before optimizer_step after gradients were reduced "
                               "early; set OVERLAP_ALLREDUCE: false when accumulating several backward passes per step")
        loss.backward()

    def _reduce_range(self, G, a, b):
        n = b - a
        if n <= 0:
            return
        per = (n + self.buckets - 1) // self.buckets
        per = (per + 63) // 64 * 64
        for i in range(a, b, per):
            dist.all_reduce(G[i:min(b, i + per)], op=dist.ReduceOp.SUM)

    def _early_reduce(self):
        """Called at the start of the last backward node (vision encoder): reduce everything but the vision range on a side
        stream while the vision backward runs (the reference's DDP overlaps bucket by bucket, ddp_accelerator.py:65)."""
        G = self._model.flat.G
        v0, v1 = self._vis
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self._comm_stream.wait_event(ready)
        with torch.cuda.stream(self._comm_stream):
            self._reduce_range(G, 0, v0)
            self._reduce_range(G, v1, self._train_end)
            done = torch.cuda.Event()
            done.record(self._comm_stream)
        self._early = done

    def all_reduce_grads(self, model):
        """SUM all-reduce of the trainable part of the flat gradient buffer in a few large NVLink messages (averaging is
        folded into the optimizer kernel's grad_mul).  When the non-vision ranges were already reduced under the vision
        backward (_early_reduce) only the vision range is left."""
        if self.world == 1:
            return
        m = model.module if hasattr(model, "module") else model
        G = m.flat.G
        if not hasattr(self, "_train_end"):
            self._layout(m)
        if self._early is not None:
            torch.cuda.current_stream().wait_event(self._early)
            self._early = None
            self._reduce_range(G, self._vis[0], self._vis[1])
        else:
            self._reduce_range(G, 0, self._train_end)

    def optimizer_step(self, optimizer, model):
        """clip_grad_norm_(CLIP_GRAD_NORM) + AdamW step + zero_grad (ddp_accelerator.py:89-98).  Returns the total gradient
        norm as a 1-element device tensor (no host sync; float() it if a python number is needed)."""
        self.all_reduce_grads(model)
        norm = optimizer.step(max_grad_norm=self.clip, grad_mul=1.0 / self.world)
        optimizer.zero_grad()
        return norm
