"""CPU oracle for the drvish per-minibatch ELBO forward/backward.

TEST INFRASTRUCTURE ONLY.  Nothing in ``drvish_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs do, and there only as the checker or the timed CPU baseline.

What it is
----------
A plain PyTorch (CPU, fp32 or fp64, autograd) restatement of what one
``pyro.infer.SVI.step`` computes for ``drvish.models.NBVAE`` / ``DRNBVAE``:

* networks: ``make_fc`` / ``Encoder`` / ``NBDecoder`` / ``logit_mean_sigmoid`` follow
  ``src/drvish/models/modules.py:15-23, 31-60, 63-98, 159-180`` of the reference.
* ELBO assembly: follows ``src/drvish/models/nbvae.py:57-90`` and
  ``src/drvish/models/drnbvae.py:87-132`` with Pyro's ``Trace_ELBO`` semantics for fully
  reparameterised guides.  The distributions are torch's own
  (``torch.distributions.{Normal,Laplace,NegativeBinomial}``), exactly the classes
  ``pyro.distributions`` wraps.

Parity pinning status
---------------------
* a1-a3, a7 (networks, dose-response mean): PINNED against the reference's own code.  The
  reference's ``modules.py`` is executed in this container through a small ``pyro`` import
  shim (``load_reference_modules``) and its outputs are committed as golden vectors under
  ``tests/golden/`` (generator: ``tests/golden/make_golden.py``).
* a6 (NB log-prob): calls ``torch.distributions.NegativeBinomial.log_prob`` itself.
* a4, a5, a8, a9 (Pyro's assembly of the ELBO, the dose-response prior/guide/factor terms):
  **parity unpinned** - pyro-ppl is an un-vendored, unpinned dependency of the reference
  (``setup.py:33``), absent from this image with no wheel available, and the reference
  ships no tests or golden values.  The restatement follows Pyro's published ``Trace_ELBO``
  algorithm: loss = -(sum of model log-probs - sum of guide log-probs), each scaled by the
  enclosing ``poutine.scale``, one particle, pathwise gradients.

Randomness is explicit: every noise tensor is an input, so that oracle and CUDA path see
identical draws (SURVEY.md section 8b "randomness").
"""
from __future__ import annotations

import importlib.util
import math
import os
import sys
import types
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.distributions import Laplace, NegativeBinomial, Normal

REFERENCE_MODULES_PY = "/root/reference/src/drvish/models/modules.py"


# --------------------------------------------------------------------------------------
# networks (modules.py:15-98)
# --------------------------------------------------------------------------------------
def make_fc(dims: Sequence[int]) -> nn.Sequential:
    """BatchNorm1d(in) -> ReLU -> Linear(in, out) for every consecutive pair of ``dims``,
    *including before the first Linear* (modules.py:15-23)."""
    mods: List[nn.Module] = []
    for d_in, d_out in zip(dims[:-1], dims[1:]):
        mods += [nn.BatchNorm1d(d_in), nn.ReLU(), nn.Linear(d_in, d_out)]
    return nn.Sequential(*mods)


def split_in_half(t: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """First half / second half of the last axis (modules.py:27-28)."""
    n = t.shape[-1] // 2
    return t[..., :n], t[..., n:]


class OracleEncoder(nn.Module):
    """modules.py:31-60.  Output column order of the last Linear: [z_loc(L), l_loc(1),
    z_raw_scale(L), l_raw_scale(1)]."""

    def __init__(self, n_input: int, n_output: int, layers: Sequence[int]):
        super().__init__()
        self.fc = make_fc([n_input] + list(layers) + [2 * n_output + 2])

    def forward(self, x):
        q_m, q_v = split_in_half(self.fc(torch.sqrt(x)))
        return q_m[..., :-1], F.softplus(q_v[..., :-1]), q_m[..., -1:], F.softplus(q_v[..., -1:])


class OracleNBDecoder(nn.Module):
    """modules.py:63-98."""

    def __init__(self, n_input: int, n_output: int, layers: Sequence[int]):
        super().__init__()
        self.fc = make_fc([n_input] + list(layers) + [2 * n_output])

    def forward(self, z, library):
        scale, log_r = split_in_half(self.fc(z))
        scale = F.log_softmax(scale, dim=-1)
        return log_r, library + scale - log_r


def logit_mean_sigmoid(x, weights, bias, labels):
    """modules.py:159-180: sigmoid(x @ W + bias) -> per-class mean over the minibatch ->
    logit(., eps=1e-5).  Returns (n_classes_present, n_conditions, n_targets)."""
    response = torch.sigmoid((x @ weights)[..., None, :] + bias)  # (B, C, T)
    n_cls = int(labels.max().item()) + 1
    counts = torch.bincount(labels, minlength=n_cls)
    if bool((counts == 0).any()):
        # the reference's scatter_add_ into a `unique`-sized tensor raises when a class is
        # missing from the batch
        raise RuntimeError("index out of bounds: every class 0..K-1 must occur in the batch")
    acc = torch.zeros((n_cls,) + tuple(response.shape[1:]), dtype=response.dtype)
    acc = acc.index_add(0, labels, response)
    return torch.logit(acc / counts.to(response.dtype).view(-1, 1, 1), eps=1e-5)


class OracleLMB(nn.Module):
    """Parameter holder mirroring ``LinearMultiBias`` (modules.py:101-157).  ``*_scale`` are
    PyroParams with a positive constraint, i.e. stored as ``log(scale)`` and named
    ``*_scale_unconstrained`` in the state dict."""

    def __init__(self, n_input: int, n_targets: int, n_conditions: int, lam_scale: float,
                 bias_scale: float, alpha: float = 1.0):
        super().__init__()
        self.alpha = alpha
        self.lam_scale = float(lam_scale)
        self.prior_bias_scale = float(bias_scale)
        self.weight_loc = nn.Parameter(nn.init.xavier_uniform_(torch.empty(n_input, n_targets)))
        self.weight_scale_unconstrained = nn.Parameter(torch.zeros(n_input, n_targets))
        self.bias_loc = nn.Parameter(nn.init.xavier_uniform_(torch.empty(n_conditions, n_targets)))
        self.bias_scale_unconstrained = nn.Parameter(torch.zeros(n_conditions, n_targets))


class OracleNBVAE(nn.Module):
    """nbvae.py:28-55."""

    def __init__(self, *, n_input, n_latent=8, layers=(64, 64, 64), lib_loc=7.5, lib_scale=0.5,
                 scale_factor=1.0, epsilon=1e-6):
        super().__init__()
        layers = list(layers)
        self.encoder = OracleEncoder(n_input, n_latent, layers)
        self.decoder = OracleNBDecoder(n_latent, n_input, layers[::-1])
        self.epsilon = float(epsilon)
        self.lib_loc = float(lib_loc)
        self.lib_scale = float(lib_scale)
        self.n_latent = n_latent
        self.scale_factor = float(scale_factor)


class OracleDRNBVAE(OracleNBVAE):
    """drnbvae.py:35-84."""

    def __init__(self, *, n_input, n_classes, n_drugs, n_conditions, n_latent=8,
                 layers=(64, 64, 64), lib_loc=7.5, lib_scale=0.5, lam_scale=5.0, bias_scale=10.0,
                 sigma_scale=1.0, scale_factor=1.0, epsilon=1e-6):
        super().__init__(n_input=n_input, n_latent=n_latent, layers=layers, lib_loc=lib_loc,
                         lib_scale=lib_scale, scale_factor=scale_factor, epsilon=epsilon)
        self.lmb_list = nn.ModuleList(
            [OracleLMB(n_latent, 1, n_conditions[i], lam_scale, bias_scale) for i in range(n_drugs)]
        )
        self.sigma_scale = float(sigma_scale)
        self.n_classes = n_classes


# --------------------------------------------------------------------------------------
# ELBO (nbvae.py:57-90, drnbvae.py:87-132)
# --------------------------------------------------------------------------------------
def _vae_terms(model: OracleNBVAE, x, eps_z, eps_l):
    """Guide then model for the plate("data") block.  Returns z, l and the per-cell ELBO
    contribution (before scale_factor)."""
    one = x.new_ones(())
    # guide (nbvae.py:81-90): encoder, reparameterised draws
    z_loc, z_scale, l_loc, l_scale = model.encoder(x)
    z = z_loc + eps_z * z_scale  # Normal.rsample, torch normal.py:82-85
    l = l_loc + eps_l * l_scale
    log_q = Normal(z_loc, z_scale).log_prob(z).sum(-1) + Normal(l_loc, l_scale).log_prob(l).sum(-1)
    # model (nbvae.py:57-78)
    log_p = Normal(0 * one, x.new_ones(model.n_latent)).log_prob(z).sum(-1)
    log_p = log_p + Normal(model.lib_loc * one, model.lib_scale * x.new_ones(1)).log_prob(l).sum(-1)
    log_r, logit = model.decoder(z, l)
    nb = NegativeBinomial(total_count=torch.exp(log_r) + model.epsilon, logits=logit)
    ll = nb.log_prob(x).sum(-1)
    return z, l, log_p, ll, log_q


def nbvae_loss(model: OracleNBVAE, x, eps_z, eps_l) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """loss = -scale_factor * sum_b (log p_z + log p_l + LL - log q_z - log q_l)."""
    z, l, log_p, ll, log_q = _vae_terms(model, x, eps_z, eps_l)
    elbo = model.scale_factor * (log_p + ll - log_q).sum()
    return -elbo, {"z": z, "l": l, "ll": ll, "log_p": log_p, "log_q": log_q}


def mcv_nbvae_loss(model: OracleNBVAE, x0, x1, log_split, log_split_complement, eps_z, eps_l):
    """MCVNBVAE step (reference mcv_nbvae.py:57-105): the guide encodes x0 (:98), the model scores x1
    (:81-87) with ``log_r += log_data_split_complement - log_data_split`` (:78-79) - the logits the
    decoder returned are NOT recomputed after the shift.  Same ELBO assembly as nbvae_loss."""
    one = x0.new_ones(())
    z_loc, z_scale, l_loc, l_scale = model.encoder(x0)
    z = z_loc + eps_z * z_scale
    l = l_loc + eps_l * l_scale
    log_q = Normal(z_loc, z_scale).log_prob(z).sum(-1) + Normal(l_loc, l_scale).log_prob(l).sum(-1)
    log_p = Normal(0 * one, x0.new_ones(model.n_latent)).log_prob(z).sum(-1)
    log_p = log_p + Normal(model.lib_loc * one, model.lib_scale * x0.new_ones(1)).log_prob(l).sum(-1)
    log_r, logit = model.decoder(z, l)
    log_r = log_r + (log_split_complement - log_split)
    nb = NegativeBinomial(total_count=torch.exp(log_r) + model.epsilon, logits=logit)
    ll = nb.log_prob(x1).sum(-1)
    elbo = model.scale_factor * (log_p + ll - log_q).sum()
    return -elbo, {"z": z, "l": l, "ll": ll, "log_p": log_p, "log_q": log_q}


def split_molecules(umis, data_split, data_split_complement, overlap_factor):
    """Restatement of the reference's ``MCVDataLoader.split_molecules`` (train/mcv_train.py:20-40):
    two dependent binomial thinnings and the overlap bookkeeping, with torch.distributions on the CPU.
    PINNED: bit-identical to the reference's own function under the same ``torch.manual_seed``
    (tests/golden/ref_split_molecules.npz, generated by executing the reference file).  The CUDA kernel
    draws from Philox instead, so it is compared with this function in DISTRIBUTION (pmf chi-square,
    moments) and through the exact invariants x0 = n - y_data, x1 = n - x_data."""
    x_data = torch.distributions.Binomial(umis, probs=data_split - overlap_factor).sample()
    y_data = torch.distributions.Binomial(
        umis - x_data, probs=(1 - data_split) / (1 - data_split + overlap_factor)).sample()
    overlap = umis - x_data - y_data
    return x_data + overlap, y_data + overlap, torch.log(data_split), torch.log(data_split_complement)


def drnbvae_loss(model: OracleDRNBVAE, x, labels, y: Sequence[torch.Tensor], eps_z, eps_l,
                 w_prior: Sequence[torch.Tensor], b_prior: Sequence[torch.Tensor],
                 eps_bias: Sequence[torch.Tensor], include_guide_factor: bool = True):
    """DRNBVAE step.

    ``w_prior[i]`` (L,1) / ``b_prior[i]`` (C_i,1): the draws of the model-side sites
    ``i.weight_p`` / ``i.bias_p`` from their priors (modules.py:130-135; the guide has no
    such sites so Pyro samples them from the prior inside the model).
    ``eps_bias[i]`` (C_i,1): standard-normal noise of the guide-side draw
    ``bias = bias_loc + bias_scale*eps`` feeding the monotonic factor (modules.py:189-196);
    the pathwise gradient is used for it (has_rsample=True semantics).
    The drs/prior/factor terms sit outside the plate: not multiplied by scale_factor.
    """
    z, l, log_p, ll, log_q = _vae_terms(model, x, eps_z, eps_l)
    elbo = model.scale_factor * (log_p + ll - log_q).sum()
    one = x.new_ones(())
    terms = {"z": z, "l": l, "ll": ll, "log_p": log_p, "log_q": log_q, "m": []}
    drs_total = x.new_zeros(())
    prior_total = x.new_zeros(())
    factor_total = x.new_zeros(())
    for i, lmb in enumerate(model.lmb_list):
        w, b = w_prior[i].to(x.dtype), b_prior[i].to(x.dtype)
        prior_total = prior_total + Laplace(0 * one, lmb.lam_scale * one).log_prob(w).sum()
        prior_total = prior_total + Normal(0 * one, lmb.prior_bias_scale * one).log_prob(b).sum()
        m = logit_mean_sigmoid(z, w, b, labels)  # (K, C_i, 1)
        terms["m"].append(m)
        drs_total = drs_total + Normal(m, model.sigma_scale * one).log_prob(y[i].to(x.dtype)).sum()
        if include_guide_factor:
            bias = lmb.bias_loc + torch.exp(lmb.bias_scale_unconstrained) * eps_bias[i].to(x.dtype)
            log_factor = -lmb.alpha * torch.clamp(bias[:-1, :] - bias[1:, :], min=0).sum()
            # a factor site in the GUIDE enters the ELBO with a minus sign
            factor_total = factor_total + log_factor
    elbo = elbo + drs_total + prior_total - factor_total
    terms.update(drs=drs_total, prior=prior_total, factor=factor_total)
    return -elbo, terms


def loss_and_grads(model: nn.Module, loss_fn, *args, **kwargs):
    """Run loss_fn, backward, and return (float loss, {param_name: grad}) like
    ``SVI.step`` leaves them in ``.grad``."""
    for p in model.parameters():
        p.grad = None
    loss, terms = loss_fn(model, *args, **kwargs)
    loss.backward()
    grads = {}
    for n, p in model.named_parameters():
        grads[n] = torch.zeros_like(p) if p.grad is None else p.grad.detach().clone()
    return float(loss.detach()), grads, terms


# --------------------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------------------
def synthetic_counts(n_cells: int, n_genes: int, seed: int, device="cpu") -> torch.Tensor:
    """x[b,g] ~ NB(total_count=2, logits=u_b+v_g), u_b~N(0,.5^2), v_g~N(-1,1); float32."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    u = 0.5 * torch.randn(n_cells, 1, generator=g)
    v = -1.0 + torch.randn(1, n_genes, generator=g)
    logits = u + v
    # NB(r, p) as Gamma-Poisson mixture: rate ~ Gamma(r, (1-p)/p), p = sigmoid(logits)
    rate = torch._standard_gamma(torch.full_like(logits, 2.0), generator=g) * torch.exp(logits)
    x = torch.poisson(rate, generator=g)
    return x.to(torch.float32).to(device)


def make_noise(model: OracleNBVAE, batch: int, seed: int = 1, dtype=torch.float32):
    """Noise tensors in the reference's draw order."""
    g = torch.Generator().manual_seed(seed)
    out = {
        "eps_z": torch.randn(batch, model.n_latent, generator=g, dtype=dtype),
        "eps_l": torch.randn(batch, 1, generator=g, dtype=dtype),
    }
    if isinstance(model, OracleDRNBVAE):
        w_prior, b_prior, eps_bias = [], [], []
        for lmb in model.lmb_list:
            L, C = lmb.weight_loc.shape[0], lmb.bias_loc.shape[0]
            u = torch.rand(L, 1, generator=g, dtype=dtype) * 2 - 1  # Laplace via uniform, laplace.py:73-85
            w_prior.append(-lmb.lam_scale * u.sign() * torch.log1p(-u.abs().clamp(max=1 - 1e-6)))
            b_prior.append(lmb.prior_bias_scale * torch.randn(C, 1, generator=g, dtype=dtype))
            eps_bias.append(torch.randn(C, 1, generator=g, dtype=dtype))
        out.update(w_prior=w_prior, b_prior=b_prior, eps_bias=eps_bias)
    return out


CONFIGS = {
    # SURVEY 8d / BASELINE.md section 4
    1: dict(n_input=128, n_latent=16, layers=[256, 256], batch=128, drugs=2, doses=8, classes=4,
            lam_scale=1.0, bias_scale=1.0),
    2: dict(n_input=2000, n_latent=10, layers=[128, 128], batch=4096, drugs=0, doses=0, classes=0),
    3: dict(n_input=5000, n_latent=10, layers=[256, 256], batch=4096, drugs=8, doses=8, classes=8),
    4: dict(n_input=20000, n_latent=32, layers=[512, 512], batch=4096, drugs=0, doses=0, classes=0),
    5: dict(n_input=30000, n_latent=64, layers=[512, 512], batch=4096, drugs=96, doses=10, classes=8),
}


def build_model(cfg: dict, seed: int = 0) -> OracleNBVAE:
    torch.manual_seed(seed)
    if cfg.get("drugs", 0) > 0:
        return OracleDRNBVAE(n_input=cfg["n_input"], n_classes=cfg["classes"], n_drugs=cfg["drugs"],
                             n_conditions=[cfg["doses"]] * cfg["drugs"], n_latent=cfg["n_latent"],
                             layers=cfg["layers"], lam_scale=cfg.get("lam_scale", 5.0),
                             bias_scale=cfg.get("bias_scale", 10.0))
    return OracleNBVAE(n_input=cfg["n_input"], n_latent=cfg["n_latent"], layers=cfg["layers"])


def step(model, x, noise, labels=None, y=None):
    """One oracle SVI step: returns (loss, grads, terms)."""
    if isinstance(model, OracleDRNBVAE):
        return loss_and_grads(model, drnbvae_loss, x, labels, y, noise["eps_z"], noise["eps_l"],
                              noise["w_prior"], noise["b_prior"], noise["eps_bias"])
    return loss_and_grads(model, nbvae_loss, x, noise["eps_z"], noise["eps_l"])


# --------------------------------------------------------------------------------------
# executing the reference's own modules.py (validation + golden generation; never at
# run time on the GPU box, where /root/reference does not exist)
# --------------------------------------------------------------------------------------
def load_reference_modules(path: str = REFERENCE_MODULES_PY):
    """Load the reference ``modules.py`` by file path with a minimal ``pyro`` stand-in so its
    top-level imports and the ``LinearMultiBias`` class body execute.  ``make_fc``,
    ``Encoder``, ``NBDecoder`` and the static ``logit_mean_sigmoid`` then run unmodified."""
    if not os.path.exists(path):
        return None
    saved = {k: sys.modules.get(k) for k in ("pyro", "pyro.distributions", "pyro.nn")}
    pyro = types.ModuleType("pyro")
    pdist = types.ModuleType("pyro.distributions")
    pnn = types.ModuleType("pyro.nn")

    class _Dist:
        def __init__(self, *a, **k):
            pass

        def expand(self, *a, **k):
            return self

        def to_event(self, *a, **k):
            return self

    pdist.Laplace = pdist.Normal = pdist.NegativeBinomial = _Dist
    pnn.PyroModule = nn.Module
    pnn.PyroParam = lambda value, constraint=None: nn.Parameter(value)
    pnn.PyroSample = lambda prior: None
    pyro.distributions, pyro.nn = pdist, pnn
    pyro.factor = lambda *a, **k: None
    sys.modules.update({"pyro": pyro, "pyro.distributions": pdist, "pyro.nn": pnn})
    try:
        spec = importlib.util.spec_from_file_location("_drvish_ref_modules", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
