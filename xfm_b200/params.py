"""Flat parameter storage for the B200 path.

All parameters of a model live in ONE fp32 buffer `P` (master weights, what state_dict / the optimizer see),
with a same-layout fp32 gradient buffer `G` and a bf16 shadow `S` that the tcgen05 GEMMs read.  The
`nn.Parameter`s handed to PyTorch are views into `P` (their `.grad` views into `G`), so

  * the bf16 shadow is refreshed by one cast kernel over the whole buffer after an optimizer step,
  * data-parallel gradient reduction is a handful of large NCCL all-reduces over slices of `G`
    (replacing the reference's per-bucket DDP path, accelerators/ddp_accelerator.py:65),
  * parameters that are separate tensors in the reference (query/key/value of xroberta.py:170-176) can be laid
    out back to back and consumed as one fused [3D, D] operand without copies.

Every segment starts on a 64-element boundary so bf16 views satisfy the 16-byte TMA alignment and fp32 views
allow 128-bit vector access.
"""
from collections import OrderedDict

import torch

from . import lib as L

ALIGN = 64


class Segment:
    __slots__ = ("name", "shape", "offset", "numel", "trainable")

    def __init__(self, name, shape, offset, numel, trainable):
        self.name, self.shape, self.offset, self.numel, self.trainable = name, tuple(shape), offset, numel, trainable


class FlatParams:
    def __init__(self):
        self.segments = OrderedDict()
        self._size = 0
        self._init = {}
        self.P = self.G = self.S = None
        self._shadow_version = -1
        self.touched = set()

    # ---- layout -------------------------------------------------------------------------------------
    def add(self, name, shape, init=None, trainable=True):
        assert self.P is None and name not in self.segments, name
        numel = 1
        for d in shape:
            numel *= d
        self.segments[name] = Segment(name, shape, self._size, numel, trainable)
        if init is not None:
            self._init[name] = init
        self._size += (numel + ALIGN - 1) // ALIGN * ALIGN
        return name

    def finalize(self, device):
        self.P = torch.zeros(self._size, dtype=torch.float32, device=device)
        for name, init in self._init.items():
            self.view32(name).copy_(init.to(device=device, dtype=torch.float32).reshape(self.segments[name].shape))
        self._init = {}
        self._alloc_side_buffers()

    def _alloc_side_buffers(self):
        self.G = torch.zeros_like(self.P)
        self.S = torch.empty(self._size, dtype=torch.bfloat16, device=self.P.device)
        self._shadow_version = -1

    def move(self, fn):
        """Apply an nn.Module._apply function (e.g. .cuda()) to the master buffer; rebuild G / S."""
        newP = fn(self.P)
        if newP.dtype != torch.float32:
            raise RuntimeError("xfm_b200 keeps fp32 master weights; .half()/.bfloat16() on the module is not supported "
                               "(compute already runs in bf16)")
        moved = newP.device != self.P.device or newP.data_ptr() != self.P.data_ptr()
        self.P = newP
        if moved:
            self._alloc_side_buffers()
        return moved

    # ---- views --------------------------------------------------------------------------------------
    def _view(self, buf, name):
        s = self.segments[name]
        return buf[s.offset:s.offset + s.numel].view(s.shape)

    def view32(self, name):
        return self._view(self.P, name)

    def view16(self, name):
        return self._view(self.S, name)

    def grad(self, name):
        self.touched.add(name)
        return self._view(self.G, name)

    def grad_padded(self, name, n):
        """1-D gradient view of `n` >= numel elements starting at the segment (the tail is alignment padding that
        nothing reads): lets column-sum kernels that need a multiple-of-8 width write odd-sized bias gradients."""
        s = self.segments[name]
        assert n >= s.numel and n <= (s.numel + ALIGN - 1) // ALIGN * ALIGN, (name, n)
        self.touched.add(name)
        return self.G[s.offset:s.offset + n]

    def _span(self, buf, first, last, shape):
        a, b = self.segments[first], self.segments[last]
        n = 1
        for d in shape:
            n *= d
        assert b.offset + b.numel - a.offset == n, (first, last, shape)  # contiguous, unpadded run
        return buf[a.offset:a.offset + n].view(shape)

    def span32(self, first, last, shape):
        return self._span(self.P, first, last, shape)

    def span16(self, first, last, shape):
        return self._span(self.S, first, last, shape)

    def span_grad(self, names, shape):
        for n in names:
            self.touched.add(n)
        return self._span(self.G, names[0], names[-1], shape)

    # ---- maintenance --------------------------------------------------------------------------------
    def sync_shadow(self, force=False):
        """bf16 shadow <- fp32 master, if the master changed (optimizer step, load_state_dict, manual edit)."""
        if not self.P.is_cuda:   # model still on the host (constructed before .to(device)): cast on first use on the GPU
            self._shadow_version = -1
            return
        v = self.P._version
        if force or v != self._shadow_version:
            L.cast_to_bf16(self.P, self.S)
            self._shadow_version = v

    def train_end(self):
        """End (aligned) of the last trainable segment: nothing behind it (the frozen VQ-KD tokenizer, 437 MB) ever
        receives a gradient, so zero_grad and the all-reduce stop there."""
        if getattr(self, "_train_end_cache", None) is None or self._train_end_cache[0] != len(self.segments):
            end = max([s.offset + (s.numel + ALIGN - 1) // ALIGN * ALIGN for s in self.segments.values() if s.trainable] or [0])
            self._train_end_cache = (len(self.segments), end)
        return self._train_end_cache[1]

    def zero_grad(self):
        self.G[:self.train_end()].zero_()
        self.touched.clear()
