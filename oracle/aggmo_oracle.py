"""CPU restatement of the reference optimizer step - TEST INFRASTRUCTURE ONLY (never imported by the
product path).

Follows ``AggMo.step`` (reference src/drvish/train/aggmo.py:42-79; the learning rate is divided by
``len(betas)`` at construction, :22) preceded by the per-parameter gradient clipping that Pyro's
``PyroOptim`` applies when built with ``clip_args={"clip_norm": c}`` (reference notebook cell 10;
``torch.nn.utils.clip_grad_norm_(p, c)`` on EACH parameter separately: coef = c / (norm + 1e-6),
clamped to 1).  PINNED against the reference's own aggmo.py executed in the build container:
tests/golden/ref_aggmo.npz (generator: tests/golden/make_golden.py:ref_aggmo)."""
from typing import Dict, List, Optional, Sequence

import torch


class AggMoOracle:
    def __init__(self, params: List[torch.Tensor], lr: float, betas: Sequence[float] = (0.0, 0.9, 0.99),
                 weight_decay: float = 0.0, nesterov: bool = False, clip_norm: Optional[float] = None):
        self.params = params
        self.lr = lr / len(betas)  # aggmo.py:22
        self.betas = list(betas)
        self.weight_decay = weight_decay
        self.nesterov = nesterov
        self.clip_norm = clip_norm
        self.bufs: List[Dict[float, torch.Tensor]] = [{b: torch.zeros_like(p) for b in self.betas} for p in params]

    @torch.no_grad()
    def step(self, grads: List[torch.Tensor]) -> None:
        for p, g, bufs in zip(self.params, grads, self.bufs):
            d_p = g.clone()
            if self.clip_norm is not None:  # torch.nn.utils.clip_grad_norm_ on this parameter alone
                total = torch.linalg.vector_norm(d_p, 2.0)
                coef = torch.clamp(self.clip_norm / (total + 1e-6), max=1.0)
                d_p = d_p * coef
            if self.weight_decay != 0:
                d_p = d_p.add(p, alpha=self.weight_decay)  # aggmo.py:63-64
            for beta in self.betas:  # aggmo.py:72-78
                buf = bufs[beta]
                buf.mul_(beta).add_(d_p)
                upd = d_p.add(buf, alpha=beta) if self.nesterov else buf
                p.add_(upd, alpha=-self.lr)
