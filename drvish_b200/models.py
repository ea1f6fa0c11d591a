"""Drop-in ``NBVAE`` / ``DRNBVAE`` (reference src/drvish/models/nbvae.py, drnbvae.py).

Constructor keywords, attributes, ``state_dict`` keys and the ``model`` / ``guide`` signatures
are the reference's.  ``model`` / ``guide`` declare the same Pyro sample sites when Pyro is
installed (their network calls are forward-only kernels); the ELBO and its gradients are
computed by :class:`drvish_b200.FusedTraceELBO`, passed as ``loss=`` to ``pyro.infer.SVI``.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from .modules import Encoder, LinearMultiBias, NBDecoder
from .runtime import ModelRuntime

try:  # Pyro is an optional dependency here (absent from the build image)
    import pyro
    import pyro.distributions as dist
    from pyro import poutine

    HAVE_PYRO = True
except Exception:  # pragma: no cover
    pyro = dist = poutine = None
    HAVE_PYRO = False


def _need_pyro():
    if not HAVE_PYRO:
        raise RuntimeError("pyro-ppl is not installed: model()/guide() need Pyro for tracing; the fused path "
                           "(drvish_b200.FusedTraceELBO().loss_and_grads(model.model, model.guide, *batch)) does not.")


class NBVAE(nn.Module):
    r"""Variational auto-encoder with a negative-binomial likelihood (reference nbvae.py:14-55).

    :param n_input: number of genes
    :param n_latent: dimensionality of the latent space
    :param layers: hidden layer widths of the encoder (the decoder gets them reversed)
    :param lib_loc, lib_scale: Normal prior of the log library size
    :param scale_factor: ``poutine.scale`` factor applied inside the data plate
    :param epsilon: added to ``exp(log_r)``
    :param device: if not None, parameters are moved there (reference: ``self.cuda(device)``)
    """

    pyro_name = "nbvae"

    def __init__(self, *, n_input: int, n_latent: int = 8, layers: Sequence[int] = (64, 64, 64), lib_loc: float = 7.5,
                 lib_scale: float = 0.5, scale_factor: float = 1.0, epsilon: float = 1e-6, device: torch.device = None):
        super().__init__()
        layers = list(layers)
        self.encoder = Encoder(n_input, n_latent, layers)
        self.decoder = NBDecoder(n_latent, n_input, layers[::-1])
        self.epsilon = torch.tensor(epsilon).to(device)
        self.l_loc = torch.tensor(lib_loc).to(device)
        self.l_scale = lib_scale * torch.ones(1, device=device)
        self.n_latent = n_latent
        self.scale_factor = scale_factor
        self._finish_init(device)

    def _finish_init(self, device) -> None:
        if device is not None:
            self.cuda(device)
        self.encoder._set_owner(self)
        self.decoder._set_owner(self)
        for lmb in getattr(self, "lmb_list", []):
            lmb._set_owner(self)
        object.__setattr__(self, "_rt", None)

    def _runtime(self) -> ModelRuntime:
        rt = self.__dict__.get("_rt")
        if rt is None:
            rt = ModelRuntime(self)
            object.__setattr__(self, "_rt", rt)
        return rt

    # -- Pyro surface (reference nbvae.py:57-90) ---------------------------------------------
    def model(self, x: torch.Tensor):
        _need_pyro()
        pyro.module(self.pyro_name, self)
        with pyro.plate("data", len(x)), poutine.scale(scale=self.scale_factor):
            z = pyro.sample("latent", dist.Normal(0, x.new_ones(self.n_latent)).to_event(1))
            l = pyro.sample("library", dist.Normal(self.l_loc, self.l_scale).to_event(1))
            log_r, logit = self.decoder(z, l)
            pyro.sample("obs", dist.NegativeBinomial(total_count=torch.exp(log_r) + self.epsilon,
                                                     logits=logit).to_event(1), obs=x)

    def guide(self, x: torch.Tensor):
        _need_pyro()
        pyro.module(self.pyro_name, self)
        with pyro.plate("data", len(x)), poutine.scale(scale=self.scale_factor):
            z_loc, z_scale, l_loc, l_scale = self.encoder(x)
            pyro.sample("latent", dist.Normal(z_loc, z_scale).to_event(1))
            pyro.sample("library", dist.Normal(l_loc, l_scale).to_event(1))


class MCVNBVAE(NBVAE):
    r"""Molecular-cross-validation variant (reference src/drvish/models/mcv_nbvae.py:14-105): the guide
    encodes one half of the molecules (``x0``), the model scores the other half (``x1``) with ``log_r``
    shifted by ``log_data_split_complement - log_data_split`` (:78-79).  Same constructor as ``NBVAE``;
    batches are ``(x0, x1, log_data_split, log_data_split_complement)`` as produced by the reference's
    ``MCVDataLoader.split_molecules`` (train/mcv_train.py:20-40)."""

    pyro_name = "mcv_nbvae"
    is_mcv = True

    def model(self, x0: torch.Tensor, x1: torch.Tensor, log_data_split: torch.Tensor,
              log_data_split_complement: torch.Tensor):
        _need_pyro()
        pyro.module(self.pyro_name, self)
        with pyro.plate("data", len(x0)), poutine.scale(scale=self.scale_factor):
            z = pyro.sample("latent", dist.Normal(0, x0.new_ones(self.n_latent)).to_event(1))
            l = pyro.sample("library", dist.Normal(self.l_loc, self.l_scale).to_event(1))
            log_r, logit = self.decoder(z, l)
            log_r = log_r + (log_data_split_complement - log_data_split)  # adjust for data split
            pyro.sample("obs", dist.NegativeBinomial(total_count=torch.exp(log_r) + self.epsilon,
                                                     logits=logit).to_event(1), obs=x1)

    def guide(self, x0: torch.Tensor, x1: torch.Tensor, log_data_split: torch.Tensor,
              log_data_split_complement: torch.Tensor):
        _need_pyro()
        pyro.module(self.pyro_name, self)
        with pyro.plate("data", len(x0)), poutine.scale(scale=self.scale_factor):
            z_loc, z_scale, l_loc, l_scale = self.encoder(x0)
            pyro.sample("latent", dist.Normal(z_loc, z_scale).to_event(1))
            pyro.sample("library", dist.Normal(l_loc, l_scale).to_event(1))


class DRNBVAE(NBVAE):
    r"""NB-VAE with a per-drug dose-response likelihood (reference drnbvae.py:16-132).

    Extra parameters: ``n_classes`` (cell classes), ``n_drugs``, ``n_conditions`` (doses per
    drug), ``lam_scale`` / ``bias_scale`` (priors of the regression weights / biases),
    ``sigma_scale`` (observation noise of the bulk dose-response data).
    """

    pyro_name = "drnbvae"

    def __init__(self, *, n_input: int, n_classes: int, n_drugs: int, n_conditions: Sequence[int], n_latent: int = 8,
                 layers: Sequence[int] = (64, 64, 64), lib_loc: float = 7.5, lib_scale: float = 0.5,
                 lam_scale: float = 5.0, bias_scale: float = 10.0, sigma_scale: float = 1.0, scale_factor: float = 1.0,
                 epsilon: float = 1e-6, device: torch.device = None):
        nn.Module.__init__(self)
        layers = list(layers)
        self.encoder = Encoder(n_input, n_latent, layers)
        self.decoder = NBDecoder(n_latent, n_input, layers[::-1])
        self.lmb_list = nn.ModuleList(
            [LinearMultiBias(n_latent, 1, n_conditions[i], lam_scale, bias_scale) for i in range(n_drugs)])
        self.epsilon = torch.tensor(epsilon).to(device)
        self.sigma_scale = torch.tensor(sigma_scale).to(device)
        self.l_loc = torch.tensor(lib_loc).to(device)
        self.l_scale = lib_scale * torch.ones(1, device=device)
        self.n_latent = n_latent
        self.n_classes = n_classes
        self.scale_factor = scale_factor
        self._finish_init(device)

    def model(self, x: torch.Tensor, labels: torch.Tensor, y: List[torch.Tensor]):
        _need_pyro()
        pyro.module(self.pyro_name, self)
        with pyro.plate("data", len(x)), poutine.scale(scale=self.scale_factor):
            z = pyro.sample("latent", dist.Normal(0, x.new_ones(self.n_latent)).to_event(1))
            l = pyro.sample("library", dist.Normal(self.l_loc, self.l_scale).to_event(1))
            log_r, logit = self.decoder(z, l)
            pyro.sample("obs", dist.NegativeBinomial(total_count=torch.exp(log_r) + self.epsilon,
                                                     logits=logit).to_event(1), obs=x)
        for i, lmb in enumerate(self.lmb_list):
            w = pyro.sample(f"{i}.weight_p", dist.Laplace(0.0, lmb.lam_scale).expand(list(lmb.weight_loc.shape)).to_event(2))
            b = pyro.sample(f"{i}.bias_p", dist.Normal(0.0, lmb.prior_bias_scale).expand(list(lmb.bias_loc.shape)).to_event(2))
            mean_dr_logit = lmb.logit_mean_sigmoid(z, w, b, labels)
            pyro.sample(f"drs_{i}", dist.Normal(loc=mean_dr_logit, scale=self.sigma_scale).to_event(2), obs=y[i])

    def guide(self, x: torch.Tensor, labels: torch.Tensor, y: List[torch.Tensor]):
        _need_pyro()
        pyro.module(self.pyro_name, self)
        with pyro.plate("data", len(x)), poutine.scale(scale=self.scale_factor):
            z_loc, z_scale, l_loc, l_scale = self.encoder(x)
            pyro.sample("latent", dist.Normal(z_loc, z_scale).to_event(1))
            pyro.sample("library", dist.Normal(l_loc, l_scale).to_event(1))
        for i, lmb in enumerate(self.lmb_list):
            pyro.factor(f"{i}.monotonic_bias", lmb.sample_guide())
