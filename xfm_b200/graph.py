"""Whole-step CUDA-graph capture for the B200 path.

A training step of this package is ~2000 kernel launches issued from Python (forward, backward, clip,
AdamW).  At the pre-training batch size the GPU is the bottleneck, but the fine-tuning configurations (24-32 samples per
GPU: Retrieval_coco.yaml / NLVR.yaml / VQA.yaml) are bound by the host's launch rate.  `GraphedStep` records one complete
step — forward, loss, backward, gradient clipping, optimizer update, zero_grad — into ONE CUDA graph and replays it with a
single launch per step.

What makes the step replayable (nothing in it may depend on a host value that changes between steps):
  * inputs live in static device tensors that each call overwrites (`copy_`, outside the graph);
  * dropout / DropPath / hard-negative seeds are kernel ARGUMENTS, which a graph bakes in: every such kernel adds the
    device-resident seed salt (lib.seed_salt_bump), advanced by a one-thread kernel at the end of the captured step
    (torch.rand for DropPath is handled by torch's own graph-safe generator state);
  * AdamW's per-parameter step counters are device-resident (optim.cu) and its hyper-parameters sit in a static device
    block refreshed between replays, so an LR scheduler keeps working;
  * the MIM block masks of masking_generator.py are sampled on the host (bit-exact RNG streams) into pinned
    buffers and copied into static device tensors before each replay.
Under torchrun the captured step contains the NCCL collectives as graph nodes (ITC all-gather, the overlapped gradient
all-reduces on the side stream become cross-stream graph edges); capture then uses capture_error_mode="thread_local"
because NCCL's watchdog thread queries events while the capture is open.
"""
import torch

from . import lib as L
from .masking import sample_batch


class GraphedStep:
    """step = GraphedStep(model, optimizer, accelerator, loss_fn, example_inputs); loss = step(inputs)

    loss_fn(model, inputs: dict) -> scalar loss tensor (may return a tuple whose first element is the loss; every tensor
    in it is exposed, detached, as static outputs).  `inputs` is a dict of tensors with fixed shapes / dtypes."""

    def __init__(self, model, optimizer, accelerator, loss_fn, example_inputs, warmup=3, uses_mim_masks=None):
        self.model, self.opt, self.acc, self.loss_fn = model, optimizer, accelerator, loss_fn
        core = model.module if hasattr(model, "module") and hasattr(model.module, "flat") else model
        self.core = core
        dev = core.flat.P.device
        self.static = {k: v.to(dev).clone() for k, v in example_inputs.items()}
        self.uses_masks = bool(uses_mim_masks) if uses_mim_masks is not None else False
        B = next(iter(self.static.values())).shape[0]
        self._B = B
        if self.uses_masks:
            m, rows = sample_batch(core._sampler, B)
            core._static_masks = (m.to(dev), rows.to(dev))
        optimizer.static_hparams(accelerator.clip, 1.0 / accelerator.world)
        # warm-up on a side stream (allocator, lazy initialisation, autograd streams), as torch.cuda.graphs requires
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        mode = "global" if accelerator.world == 1 else "thread_local"
        with torch.cuda.graph(self.graph, capture_error_mode=mode):
            self.outputs = self._eager_step()
            L.seed_salt_bump(1)
        self.launches_per_replay = 1

    def _eager_step(self):
        out = self.loss_fn(self.model, self.static)
        loss = out[0] if isinstance(out, (tuple, list)) else out
        self.acc.backward_step(loss, self.opt)
        norm = self.acc.optimizer_step(self.opt, self.model)
        outs = tuple(t.detach() for t in (out if isinstance(out, (tuple, list)) else (out,)))
        return outs + (norm,)

    def refill(self, inputs):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        if self.uses_masks:   # same host RNG streams and call order as the eager path (beit2.py:432-439)
            m, rows = sample_batch(self.core._sampler, self._B)
            # fresh pinned staging tensors: torch's pinned-memory allocator keeps each alive until its copy has executed
            self.core._static_masks[0].copy_(m.pin_memory(), non_blocking=True)
            self.core._static_masks[1].copy_(rows.pin_memory(), non_blocking=True)
        self.opt.refresh_hparams()

    # ---- prefetching loader pattern: the host -> device copy of step i+1 runs on a side stream while the graph of step i
    # executes; only a device -> device copy into the static tensors (microseconds) sits between two replays
    def stage(self, host_inputs):
        """Start copying a (pinned) host batch into the free staging slot, on the copy stream."""
        if not hasattr(self, "_slots"):
            dev = self.core.flat.P.device
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slots = [{k: torch.empty_like(v) for k, v in self.static.items()} for _ in range(2)]
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._consumed = [torch.cuda.Event(), torch.cuda.Event()]
            for e in self._consumed:
                e.record(torch.cuda.current_stream(dev))
            self._next_slot, self._staged = 0, None
            if self.uses_masks:
                for sl in self._slots:
                    sl["__mask"] = torch.empty_like(self.core._static_masks[0])
                    sl["__rows"] = torch.empty_like(self.core._static_masks[1])
        s = self._next_slot
        masks = None
        if self.uses_masks:   # host sampling for the NEXT step happens here, while the current replay runs on the device
            m, rows = sample_batch(self.core._sampler, self._B)
            masks = (m.pin_memory(), rows.pin_memory())
        self._copy_stream.wait_event(self._consumed[s])      # the commit that last read this slot has finished
        with torch.cuda.stream(self._copy_stream):
            for k, v in host_inputs.items():
                self._slots[s][k].copy_(v, non_blocking=True)
            if masks is not None:
                self._slots[s]["__mask"].copy_(masks[0], non_blocking=True)
                self._slots[s]["__rows"].copy_(masks[1], non_blocking=True)
            self._ready[s].record(self._copy_stream)
        self._staged, self._next_slot = s, s ^ 1

    def commit(self):
        """Move the staged batch into the static input tensors (device -> device) and refresh masks / hyper-parameters."""
        s = self._staged
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[s])
        for k, v in self._slots[s].items():
            if k == "__mask":
                self.core._static_masks[0].copy_(v, non_blocking=True)
            elif k == "__rows":
                self.core._static_masks[1].copy_(v, non_blocking=True)
            else:
                self.static[k].copy_(v, non_blocking=True)
        self._consumed[s].record(cur)
        self.opt.refresh_hparams()

    def __call__(self, inputs=None):
        """Replays the captured step on `inputs` (or on whatever the static tensors hold).  Returns the static output tensors
        (loss first, gradient norm last); their values are overwritten by the next call."""
        if inputs is not None:
            self.refill(inputs)
        self.graph.replay()
        return self.outputs
