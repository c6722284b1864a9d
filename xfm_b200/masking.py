"""Host-side block-wise MIM mask sampling (models/masking_generator.py:26-105, called once per image at
models/beit2.py:432-439).

Bit-exact contract: for equal `random.seed` / `np.random.seed` the masks equal the reference's, so the draws from the
two GLOBAL streams happen in the reference's order — per attempt: uniform(area), uniform(log-aspect), then (only when
the rectangle fits) randint(top), randint(left); the final trim / top-up uses one np.random.choice.  The rectangle fill
itself is vectorised.  Masks stay on the host until the batch is stacked; the row indices of the masked patches are
kept beside the bool tensor so the MIM head never needs a device-side nonzero().
"""
import math
import random

import numpy as np
import torch


class BlockMaskSampler:
    def __init__(self, grid, num_masking_patches, min_num_patches=4, max_num_patches=None, min_aspect=0.3, max_aspect=None):
        self.h, self.w = (grid, grid) if isinstance(grid, int) else grid
        self.target = num_masking_patches
        self.lo = min_num_patches
        self.hi = num_masking_patches if max_num_patches is None else max_num_patches
        hi_aspect = max_aspect or 1 / min_aspect
        self.log_aspect = (math.log(min_aspect), math.log(hi_aspect))

    def _try_rectangle(self, mask, budget):
        """Up to 10 proposals; paints the first one that adds between 1 and `budget` new patches."""
        for _ in range(10):
            area = random.uniform(self.lo, budget)
            aspect = math.exp(random.uniform(*self.log_aspect))
            rh = int(round(math.sqrt(area * aspect)))
            rw = int(round(math.sqrt(area / aspect)))
            if rw >= self.w or rh >= self.h:
                continue
            top = random.randint(0, self.h - rh)
            left = random.randint(0, self.w - rw)
            window = mask[top:top + rh, left:left + rw]
            fresh = rh * rw - int(window.sum())
            if 0 < fresh <= budget:
                window[...] = 1
                return fresh
        return 0

    def __call__(self):
        mask = np.zeros((self.h, self.w), dtype=np.int32)
        painted = 0
        while painted < self.target:
            added = self._try_rectangle(mask, min(self.target - painted, self.hi))
            if added == 0:
                break
            painted += added
        if painted != self.target:
            surplus = painted > self.target
            ys, xs = (mask if surplus else mask == 0).nonzero()
            pick = np.random.choice(ys.shape[0], abs(painted - self.target), replace=False)
            mask[ys[pick], xs[pick]] = 0 if surplus else 1
        return mask


def sample_batch(sampler, B):
    """B sampler calls (one per image, batch order) -> (bool [B, np] CPU tensor, int64 flat indices of masked patches)."""
    m = np.stack([sampler().reshape(-1) for _ in range(B)])
    return torch.from_numpy(m.astype(np.bool_)), torch.from_numpy(np.flatnonzero(m).astype(np.int64))
