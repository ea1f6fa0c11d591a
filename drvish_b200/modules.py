"""nn.Module mirrors of ``drvish.models.modules`` (reference src/drvish/models/modules.py).

Same constructor signatures, attribute names and ``state_dict`` keys as the reference, so
reference checkpoints load and reference-style code keeps working; the arithmetic runs in
the sm_100a kernels of ``libdrvish_b200.so``.  These classes are parameter containers plus
forward-only entry points (``Encoder.forward``, ``NBDecoder.forward``,
``LinearMultiBias.calc_response``); training goes through
:class:`drvish_b200.FusedTraceELBO`, which evaluates the whole ELBO step fused.
"""
from __future__ import annotations

import warnings
import weakref
from typing import Sequence, Tuple

import torch
import torch.nn as nn


def make_fc(dims: Sequence[int]) -> nn.Sequential:
    """BatchNorm1d -> ReLU -> Linear per consecutive pair, including before the first Linear
    (reference modules.py:15-23)."""
    layers = []
    for in_dim, out_dim in zip(dims, dims[1:]):
        layers.append(nn.BatchNorm1d(in_dim))
        layers.append(nn.ReLU())
        layers.append(nn.Linear(in_dim, out_dim))
    return nn.Sequential(*layers)


def split_in_half(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference modules.py:27-28"""
    return t.reshape(t.shape[:-1] + (2, -1)).unbind(-2)


class _Owned(nn.Module):
    """Sub-module whose kernels need the owning model's flat parameter buffer."""

    def _owner(self):
        ref = getattr(self, "_owner_ref", None)
        owner = ref() if ref is not None else None
        if owner is None:
            raise NotImplementedError(
                f"{type(self).__name__} must belong to a drvish_b200 NBVAE/DRNBVAE: the CUDA kernels read the "
                "model's flat parameter buffer (stand-alone sub-modules are not supported)."
            )
        return owner

    def _set_owner(self, owner) -> None:
        object.__setattr__(self, "_owner_ref", weakref.ref(owner))

    @staticmethod
    def _warn_no_autograd(params) -> None:
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            warnings.warn(
                "drvish_b200 module forwards are forward-only (no autograd graph). Train with "
                "pyro.infer.SVI(model, guide, optim, loss=drvish_b200.FusedTraceELBO()).",
                stacklevel=3,
            )


class Encoder(_Owned):
    """reference modules.py:31-60"""

    def __init__(self, n_input: int, n_output: int, layers: Sequence[int]):
        super().__init__()
        self.fc = make_fc([n_input] + list(layers) + [2 * n_output + 2])

    def forward(self, x: torch.Tensor):
        """sqrt(x) -> fc -> (z_loc, z_scale, l_loc, l_scale) (reference modules.py:46-60).
        BatchNorm uses batch statistics and updates its running buffers, as in the reference,
        which never calls ``.eval()``."""
        self._warn_no_autograd(self.parameters())
        return self._owner()._runtime().encoder_forward(x)


class NBDecoder(_Owned):
    """reference modules.py:63-98"""

    def __init__(self, n_input: int, n_output: int, layers: Sequence[int]):
        super().__init__()
        self.fc = make_fc([n_input] + list(layers) + [2 * n_output])

    def forward(self, z: torch.Tensor, library: torch.Tensor):
        """fc(z) -> (log_r, logit = library + log_softmax(scale) - log_r) (reference modules.py:77-98)."""
        self._warn_no_autograd(self.parameters())
        return self._owner()._runtime().decoder_forward(z, library)


class LinearMultiBias(_Owned):
    """reference modules.py:101-196.  ``weight_scale`` / ``bias_scale`` are positive-constrained
    PyroParams in the reference, i.e. stored as ``*_unconstrained = log(scale)``; the same
    storage (and state-dict key) is used here and ``weight_scale`` / ``bias_scale`` are
    exposed as read-only properties."""

    def __init__(self, n_input: int, n_targets: int, n_conditions: int, lam_scale, bias_scale, alpha: float = 1.0):
        super().__init__()
        if n_targets != 1:
            raise NotImplementedError("the drvish models always use n_targets=1 (reference drnbvae.py:60-66)")
        self.alpha = alpha
        self.lam_scale = float(lam_scale)
        self.prior_bias_scale = float(bias_scale)
        self.n_conditions = int(n_conditions)
        self.weight_loc = nn.Parameter(nn.init.xavier_uniform_(torch.empty(n_input, n_targets)))
        self.weight_scale_unconstrained = nn.Parameter(torch.zeros(n_input, n_targets))
        self.bias_loc = nn.Parameter(nn.init.xavier_uniform_(torch.empty(n_conditions, n_targets)))
        self.bias_scale_unconstrained = nn.Parameter(torch.zeros(n_conditions, n_targets))

    @property
    def weight_scale(self) -> torch.Tensor:
        return torch.exp(self.weight_scale_unconstrained)

    @property
    def bias_scale(self) -> torch.Tensor:
        return torch.exp(self.bias_scale_unconstrained)

    @staticmethod
    def logit_mean_sigmoid(x, weights, bias, labels):
        """reference modules.py:159-180, evaluated by the dose-response kernels.
        Returns (n_classes, n_conditions, 1)."""
        from .runtime import logit_mean_sigmoid as _lms

        return _lms(x, [weights], [bias], labels)[0]

    def forward(self, x, labels, weight_p=None, bias_p=None):
        """The reference draws ``weight_p`` / ``bias_p`` from their priors through Pyro
        (modules.py:130-135,182-183); without Pyro the draws are passed in or drawn here."""
        dev = x.device
        if weight_p is None:
            weight_p = torch.distributions.Laplace(0.0, self.lam_scale).sample(self.weight_loc.shape).to(dev)
        if bias_p is None:
            bias_p = torch.distributions.Normal(0.0, self.prior_bias_scale).sample(self.bias_loc.shape).to(dev)
        return self.logit_mean_sigmoid(x, weight_p, bias_p, labels)

    def calc_response(self, x, labels):
        """reference modules.py:185-187"""
        return self.logit_mean_sigmoid(x, self.weight_loc.detach(), self.bias_loc.detach(), labels)

    def sample_guide(self):
        """reference modules.py:189-196.  Under the fused ELBO this term is evaluated inside
        ``drv_elbo_step``; stand-alone it returns the log-factor for inspection."""
        bias = self.bias_loc + self.bias_scale * torch.randn_like(self.bias_loc)
        return -self.alpha * torch.clamp(bias[:-1, :] - bias[1:, :], min=0).sum()
