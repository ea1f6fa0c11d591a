"""Shared parity harness: run the CUDA path and the CPU oracle on identical inputs, weights and
noise and compare the ELBO, every parameter gradient and the BatchNorm buffers.

Tolerance protocol: norm-wise relative error per parameter tensor
against the fp64 oracle; the Linear biases that feed a BatchNorm have an identically zero true
gradient and are tested with an absolute bound scaled by the matching weight-gradient norm."""
from __future__ import annotations

import torch

from oracle import drvish_oracle as O

ZERO_GRAD_BIASES = ("encoder.fc.2.bias", "encoder.fc.5.bias", "decoder.fc.2.bias", "decoder.fc.5.bias")


def zero_grad_biases(n_hidden_layers: int):
    """Linear biases that are immediately followed by a BatchNorm (all but the last Linear of
    each stack): the BatchNorm subtracts the batch mean, so their exact gradient is 0."""
    return tuple(f"{s}.fc.{3 * k + 2}.bias" for s in ("encoder", "decoder") for k in range(n_hidden_layers))


def make_case(cfg: dict, data_seed: int = 5, noise_seed: int = 1, model_seed: int = 0, bn_affine_random: bool = False,
              log_r_shift: float = 0.0):
    """``cfg`` may carry ``scale_factor`` (the ``poutine.scale`` of the data plate, nbvae.py:60).
    ``log_r_shift`` is added to the log_r half of the last decoder bias (``decoder.fc.<last>.bias[G:]``), i.e. it
    moves the NB dispersion r = exp(log_r) + eps by e^shift (PyTorch-default weights give r ~ 1)."""
    model = O.build_model(cfg, seed=model_seed)
    if "scale_factor" in cfg:
        model.scale_factor = float(cfg["scale_factor"])
    if log_r_shift:
        with torch.no_grad():
            model.decoder.fc[-1].bias[cfg["n_input"]:] += float(log_r_shift)
    if bn_affine_random:
        g = torch.Generator().manual_seed(123)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
                m.bias.data = 0.6 * torch.rand(m.bias.shape, generator=g) - 0.3
    B = cfg["batch"]
    x = O.synthetic_counts(B, cfg["n_input"], seed=data_seed)
    noise = O.make_noise(model, B, seed=noise_seed)
    labels = y = None
    if cfg.get("drugs", 0):
        labels = torch.arange(B) % cfg["classes"]
        g = torch.Generator().manual_seed(99)
        y = [torch.randn(cfg["classes"], cfg["doses"], 1, generator=g) for _ in range(cfg["drugs"])]
    return dict(cfg=cfg, model=model, x=x, noise=noise, labels=labels, y=y)


def oracle_step(case, dtype=torch.float64):
    cfg = case["cfg"]
    m = O.build_model(cfg, seed=0).to(dtype)
    m.scale_factor = case["model"].scale_factor
    m.load_state_dict({k: (v.to(dtype) if v.is_floating_point() else v) for k, v in case["model"].state_dict().items()})
    cast = lambda v: [t.to(dtype) for t in v] if isinstance(v, list) else v.to(dtype)
    noise = {k: cast(v) for k, v in case["noise"].items()}
    y = [t.to(dtype) for t in case["y"]] if case["y"] is not None else None
    # ReLU gate margin: the smallest |pre-activation| of every BatchNorm -> ReLU, relative to the layer's rms.  The
    # gradient is DISCONTINUOUS where a pre-activation crosses 0: an implementation whose forward error exceeds the
    # margin flips that gate, and one flipped (cell, unit) moves a weight gradient by ~1/sqrt(cells x units) of its
    # norm (2e-3 at 512 x 128) whatever the size of the forward error (tools/disp_probe2.py).
    margins = []
    hooks = [mod.register_forward_pre_hook(lambda _m, inp: margins.append(
        float(inp[0].detach().abs().min() / inp[0].detach().pow(2).mean().sqrt())))
        for mod in m.modules() if isinstance(mod, torch.nn.ReLU)]
    loss, grads, terms = O.step(m, case["x"].to(dtype), noise, case["labels"], y)
    for h in hooks:
        h.remove()
    m.relu_gate_margin = min(margins) if margins else float("inf")
    return loss, grads, terms, m


def build_cuda_model(case, device):
    import drvish_b200 as D

    cfg = case["cfg"]
    kw = dict(n_input=cfg["n_input"], n_latent=cfg["n_latent"], layers=cfg["layers"],
              scale_factor=float(case["model"].scale_factor))
    if cfg.get("drugs", 0):
        model = D.DRNBVAE(n_classes=cfg["classes"], n_drugs=cfg["drugs"], n_conditions=[cfg["doses"]] * cfg["drugs"],
                          lam_scale=cfg.get("lam_scale", 5.0), bias_scale=cfg.get("bias_scale", 10.0), **kw)
    else:
        model = D.NBVAE(**kw)
    missing = model.load_state_dict(case["model"].state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return model.to(device)


def cuda_step(case, device="cuda:0", force_simt=False, want_grads=True, model=None):
    import drvish_b200 as D

    model = model if model is not None else build_cuda_model(case, device)
    elbo = D.FusedTraceELBO(force_simt=force_simt, data_parallel=False)
    elbo.set_noise(case["noise"])
    args = [case["x"].to(device)]
    if case["labels"] is not None:
        args += [case["labels"].to(device), [t.to(device) for t in case["y"]]]
    fn = elbo.loss_and_grads if want_grads else elbo.loss
    loss = fn(model.model, model.guide, *args)
    grads = {n: (p.grad.detach().double().cpu() if p.grad is not None else None) for n, p in model.named_parameters()}
    return loss, grads, model


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    return float((a.double() - ref.double()).norm() / (ref.double().norm() + 1e-300))


def compare_step(case, device="cuda:0", force_simt=False):
    loss64, grads64, _, m64 = oracle_step(case, torch.float64)
    loss32, grads32, _, _ = oracle_step(case, torch.float32)
    loss, grads, model = cuda_step(case, device, force_simt)
    rep = {"loss": loss, "loss64": loss64, "loss_rel": abs(loss - loss64) / abs(loss64),
           "loss32_rel": abs(loss32 - loss64) / abs(loss64), "grads": {}, "launches": model._runtime().last_launches,
           "path": model._runtime().last_path, "relu_gate_margin": m64.relu_gate_margin}
    worst, worst_name = 0.0, ""
    zero_biases = zero_grad_biases(len(case["cfg"]["layers"]))
    for name, g64 in grads64.items():
        g = grads[name]
        assert g is not None, name
        if name in zero_biases:
            wn = grads64[name.replace("bias", "weight")].norm()
            e = float(g.norm() / (wn + 1e-300))
            floor = float(grads32[name].double().norm() / (wn + 1e-300))
            rep["grads"][name] = ("abs/|dW|", e, floor)
            assert e <= 1e-4, (name, e)
            continue
        if float(g64.norm()) == 0.0:
            assert float(g.norm()) == 0.0, name
            rep["grads"][name] = ("zero", 0.0, 0.0)
            continue
        e = rel_err(g, g64)
        rep["grads"][name] = ("rel", e, rel_err(grads32[name], g64))
        if e > worst:
            worst, worst_name = e, name
    rep["worst_grad_rel"], rep["worst_grad_name"] = worst, worst_name
    # BatchNorm buffers after one step
    sd = model.state_dict()
    bn_worst = 0.0
    for k, v in m64.state_dict().items():
        if "running" in k:
            bn_worst = max(bn_worst, rel_err(sd[k].cpu(), v))
        elif "num_batches" in k:
            assert int(sd[k]) == int(v), k
    rep["bn_worst_rel"] = bn_worst
    return rep


def split_case():
    """Inputs of the molecule-split golden vector (tests/golden/make_golden.py, tests/test_oracle.py)."""
    g = torch.Generator().manual_seed(77)
    umis = torch.poisson(torch.rand(23, 41, generator=g) * 6.0, generator=g)
    umis[3, 5] = 900.0  # a large count
    split = 0.5 + 0.4 * torch.rand(23, 1, generator=g)
    overlap = 0.1 * torch.rand(23, 1, generator=g)
    return umis, split, 1 - split, overlap
