"""Fused optimizer step on the runtime's flat buffers.

``FusedAggMo`` mirrors the reference optimizer ``drvish.train.aggmo.AggMo`` (src/drvish/train/aggmo.py:
same constructor keywords, ``lr`` divided by ``len(betas)`` at construction :22, ``from_exp_form`` :29-40,
``step`` :42-79) plus the per-parameter gradient clipping Pyro's ``PyroOptim`` adds when it is built
with ``clip_args={"clip_norm": c}`` (reference notebook cell 10).  The reference runs ~10 elementwise
launches per parameter tensor per step; here the whole update is ``drv_aggmo_step``: 2-3 launches
over the flat parameter / gradient / momentum buffers.

It is a ``torch.optim.Optimizer`` with ONE parameter group, so the stock torch LR schedulers
(e.g. ``CosineAnnealingWarmRestarts``, what the reference's ``train_until_plateau`` uses) drive
``param_groups[0]["lr"]``.  ``PyroFusedAggMo`` is the callable Pyro's ``SVI`` expects as ``optim``;
``PyroFusedScheduler`` adds the LR-schedule surface ``train_until_plateau`` drives (``step(epoch)``,
``optim_objs[*].T_cur / T_0``)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .models import NBVAE

MAX_BETAS = 4
MAX_TENSORS = 96


class AggMoArgs(C.Structure):  # drv_aggmo
    _fields_ = [("lr", C.c_float), ("weight_decay", C.c_float), ("clip_norm", C.c_float), ("nesterov", C.c_int32),
                ("n_betas", C.c_int32), ("betas", C.c_float * MAX_BETAS)]


class FusedAggMo(torch.optim.Optimizer):
    def __init__(self, model: NBVAE, lr: float, betas: Sequence[float] = (0.0, 0.9, 0.99), weight_decay: float = 0.0,
                 nesterov: bool = False, clip_norm: Optional[float] = None):
        if not isinstance(model, NBVAE):
            raise TypeError("FusedAggMo expects a drvish_b200 NBVAE / DRNBVAE (it updates the model's flat buffers)")
        betas = [float(b) for b in betas]
        if not 1 <= len(betas) <= MAX_BETAS:
            raise ValueError(f"FusedAggMo supports 1..{MAX_BETAS} betas")
        self.model = model
        self.clip_norm = clip_norm
        defaults = dict(lr=lr / len(betas), betas=betas, weight_decay=weight_decay, nesterov=nesterov)  # aggmo.py:21-26
        super().__init__([p for p in model.parameters()], defaults)
        self._momentum = None
        self._scratch = None

    @classmethod
    def from_exp_form(cls, model: NBVAE, lr: float, a: float = 0.1, k: int = 3, weight_decay: float = 0.0,
                      nesterov: bool = False, clip_norm: Optional[float] = None):
        betas = [1 - a ** i for i in range(k)]  # aggmo.py:39
        return cls(model, lr, betas, weight_decay, nesterov, clip_norm)

    # momentum buffers: n_betas consecutive copies of the flat layout
    def _buffers(self, rt):
        n = rt.flat_params.numel()
        nb = len(self.param_groups[0]["betas"])
        if self._momentum is None or self._momentum.numel() != nb * n or self._momentum.device != rt.flat_params.device:
            self._momentum = torch.zeros(nb * n, dtype=torch.float32, device=rt.flat_params.device)
            self._scratch = torch.zeros(MAX_TENSORS, dtype=torch.float64, device=rt.flat_params.device)
        return self._momentum, self._scratch

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from .runtime import _ptr, _stream

        rt = self.model._runtime()
        rt.ensure_flat()
        if rt.flat_grads is None:
            raise RuntimeError("FusedAggMo.step: no gradients yet (run FusedTraceELBO.loss_and_grads first)")
        g = self.param_groups[0]
        mom, scratch = self._buffers(rt)
        hp = AggMoArgs()
        hp.lr, hp.weight_decay = float(g["lr"]), float(g["weight_decay"])
        hp.clip_norm = float(self.clip_norm) if self.clip_norm is not None else 0.0
        hp.nesterov, hp.n_betas = int(bool(g["nesterov"])), len(g["betas"])
        for i, b in enumerate(g["betas"]):
            hp.betas[i] = float(b)
        dims = rt.dims(2)
        _lib.check(rt.lib.drv_aggmo_step(C.byref(dims), C.byref(hp), _ptr(rt.flat_params), _ptr(rt.flat_grads), _ptr(mom),
                                         _ptr(rt.dose_offsets) if rt.n_drugs else None, _ptr(scratch),
                                         _stream(rt.device)), "drv_aggmo_step")
        return loss

    def zero_grad(self, set_to_none: bool = False):  # the ELBO step clears the gradient buffer itself
        return None

    def state_dict(self):
        sd = super().state_dict()
        sd["momentum_flat"] = None if self._momentum is None else self._momentum.clone()
        return sd

    def load_state_dict(self, sd):
        sd = dict(sd)
        mom = sd.pop("momentum_flat", None)
        super().load_state_dict(sd)
        if mom is not None:
            self._momentum = mom.clone()
            self._scratch = torch.zeros(MAX_TENSORS, dtype=torch.float64, device=mom.device)


class PyroFusedAggMo:
    """Drop-in for the reference's ``PyroAggMo(optim_args, clip_args)`` (train/__init__.py:17-21) as the
    ``optim`` argument of ``pyro.infer.SVI``: SVI calls ``optim(params)`` after ``loss_and_grads``."""

    def __init__(self, model: NBVAE, optim_args: dict, clip_args: Optional[dict] = None):
        clip = None if not clip_args else clip_args.get("clip_norm")
        self.optim = FusedAggMo(model, clip_norm=clip, **optim_args)

    def __call__(self, params, *args, **kwargs) -> None:
        self.optim.step()

    def get_state(self):
        return self.optim.state_dict()

    def set_state(self, state) -> None:
        self.optim.load_state_dict(state)


class PyroFusedScheduler:
    """Drop-in for a Pyro LR scheduler over the reference optimizer - notebook cell 10:
    ``pyro.optim.CosineAnnealingWarmRestarts({"optimizer": AggMo, "optim_args": {...}, "T_0": ..., ...}, clip_args)`` -
    i.e. everything ``pyro.infer.SVI`` and ``train_until_plateau`` (train/__init__.py:54-110) touch:

    * ``__call__(params)``: what ``SVI.step`` calls after ``loss_and_grads`` - one fused AggMo step;
    * ``step(epoch)``: what ``train_until_plateau`` calls once per epoch (:95) - advances the schedule;
    * ``optim_objs``: ``{key: scheduler}`` whose values expose ``T_cur`` / ``T_0`` (:96: "did a cosine cycle just
      end?").  Pyro keeps one optimizer + scheduler per parameter tensor, all stepped in lock-step; here there is
      ONE ``FusedAggMo`` over the flat buffers and hence one scheduler object, which answers the same question;
    * ``get_state`` / ``set_state``.

    ``scheduler_constructor`` is the torch scheduler class (default ``CosineAnnealingWarmRestarts``); ``optim_args``
    has Pyro's layout: ``{"optimizer": <ignored: the fused AggMo is used>, "optim_args": {lr, betas, ...}, **scheduler_kwargs}``."""

    def __init__(self, model: NBVAE, optim_args: dict, clip_args: Optional[dict] = None, scheduler_constructor=None):
        if scheduler_constructor is None:
            scheduler_constructor = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts
        args = dict(optim_args)
        args.pop("optimizer", None)
        inner = args.pop("optim_args", {})
        clip = None if not clip_args else clip_args.get("clip_norm")
        self.optim = FusedAggMo(model, clip_norm=clip, **inner)
        self.scheduler = scheduler_constructor(self.optim, **args)
        self.optim_objs = {"flat_buffers": self.scheduler}

    def __call__(self, params, *args, **kwargs) -> None:
        self.optim.step()

    def step(self, *args, **kwargs) -> None:
        import warnings

        with warnings.catch_warnings():  # torch deprecates the epoch argument the reference passes
            warnings.simplefilter("ignore", UserWarning)
            for sched in self.optim_objs.values():
                sched.step(*args, **kwargs)

    def get_state(self):
        return {"optimizer": self.optim.state_dict(), "scheduler": self.scheduler.state_dict()}

    def set_state(self, state) -> None:
        self.optim.load_state_dict(state["optimizer"])
        self.scheduler.load_state_dict(state["scheduler"])
