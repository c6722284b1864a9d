"""CPU oracle for the optimizer half of the hot path: parameter grouping, gradient clipping and the AdamW update.

TEST INFRASTRUCTURE ONLY (same rules as oracle/xfm_oracle.py: imported by tests/, __graft_entry__.smoke() and bench.py's
CPU legs, never by xfm_b200/).

What it restates, file:line relative to the reference root:
  * optim.py:4-50          create_optimizer: 4 groups = {decay, no-decay} x {lr, lr * lr_mult}; the no-decay rule is a
                           SUBSTRING test of the parameter name against 9 patterns; `init_params` names get lr * lr_mult
  * accelerators/ddp_accelerator.py:89-98   optimizer_step: torch.nn.utils.clip_grad_norm_(params, CLIP_GRAD_NORM),
                           optimizer.step(), optimizer.zero_grad()
  * transformers.optimization.AdamW.step — THIRD-PARTY (transformers==4.12.5, requirements.txt:2; the class no longer
    exists in the installed transformers 5.x, so it cannot be imported here).  Its published update, per parameter with a
    gradient:  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;  step = lr * sqrt(1 - b2^t) / (1 - b1^t)  (correct_bias);
    p <- p - step * m / (sqrt(v) + eps);  then, if weight_decay > 0,  p <- p - lr * wd * p  (decoupled, AFTER the Adam
    update, with the un-corrected lr).  Parameters whose grad is None are skipped (no state change).

Pin: tests/test_oracle_golden.py checks `hf_adamw_step` with weight_decay = 0 against torch.optim.Adam (identical up to the
placement of eps, i.e. to ~eps / |g|) and `clip_grad_norm` against torch.nn.utils.clip_grad_norm_ exactly.
"""
import math

import torch

NO_DECAY = ("bias", "LayerNorm.bias", "LayerNorm.weight", "norm.bias", "norm.weight", "norm1.bias", "norm1.weight",
            "norm2.bias", "norm2.weight")  # optim.py:17-25


def group_of(name, init_params):
    """optim.py:31-46: index into the 4 parameter groups for a trainable parameter called `name`."""
    no_decay = any(nd in name for nd in NO_DECAY)
    large = name in init_params
    return (1 if no_decay else 0) + (2 if large else 0)


def group_hparams(lr, weight_decay, lr_mult):
    """optim.py:10-15."""
    return [(lr, weight_decay), (lr, 0.0), (lr * lr_mult, weight_decay), (lr * lr_mult, 0.0)]


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (norm_type 2) as called at ddp_accelerator.py:94-95: returns (total_norm, clip_coef);
    the caller multiplies every gradient by clip_coef.  total_norm = || [ ||g_i|| ] ||, coef = min(1, max / (total + 1e-6))."""
    norms = [g.detach().float().norm(2) for g in grads]
    total = torch.stack(norms).norm(2) if norms else torch.tensor(0.0)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return total, coef


def hf_adamw_step(p, g, m, v, step, lr, weight_decay, betas=(0.9, 0.98), eps=1e-8, correct_bias=True):
    """One transformers-4.12.5 AdamW update of tensor p in place (see the module docstring).  step is 1-based."""
    b1, b2 = betas
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = step_size * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


class RefOptimizer:
    """clip_grad_norm_ + HF AdamW over a dict of named leaf tensors, grouped by optim.py's rules."""

    def __init__(self, named, init_params, lr=1e-4, weight_decay=0.01, lr_mult=1.0, betas=(0.9, 0.98), eps=1e-8,
                 max_grad_norm=0.0):
        self.named = dict(named)
        self.hp = group_hparams(lr, weight_decay, lr_mult)
        self.group = {n: group_of(n, set(init_params)) for n in self.named}
        self.betas, self.eps, self.max_grad_norm = betas, eps, max_grad_norm
        self.state = {}
        self.t = {}

    @torch.no_grad()
    def step(self, grad_mul=1.0):
        """Uses and clears p.grad of every tensor; returns the total gradient norm (after grad_mul, before clipping)."""
        live = [(n, p) for n, p in self.named.items() if p.grad is not None]
        grads = [p.grad * grad_mul for _, p in live]
        total, coef = torch.tensor(0.0), torch.tensor(1.0)
        if self.max_grad_norm > 0:
            total, coef = clip_grad_norm(grads, self.max_grad_norm)
        for (n, p), g in zip(live, grads):
            g = g * coef
            if n not in self.state:
                self.state[n] = (torch.zeros_like(p), torch.zeros_like(p))
                self.t[n] = 0
            self.t[n] += 1
            lr, wd = self.hp[self.group[n]]
            m, v = self.state[n]
            hf_adamw_step(p, g, m, v, self.t[n], lr, wd, self.betas, self.eps)
        for _, p in self.named.items():
            p.grad = None
        return total
