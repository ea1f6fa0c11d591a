#!/usr/bin/env python
"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

* ``ref_modules_small.npz``: reference ``Encoder`` / ``NBDecoder`` / ``logit_mean_sigmoid``
  (``/root/reference/src/drvish/models/modules.py``, loaded through the pyro import shim in
  ``oracle.drvish_oracle.load_reference_modules``) evaluated forward AND backward on seeded
  inputs, fp32 and fp64.  Pins SURVEY 8a rows a1-a3, a7.
* ``oracle_step_cfg1.npz`` / ``oracle_step_tiny.npz``: full-step loss and parameter gradients
  of the oracle (fp64) for the config-1 shape (DRNBVAE) and a ragged tiny NBVAE, so the GPU
  box can check the CUDA path without /root/reference.  These are ORACLE outputs (the ELBO
  assembly is unpinned by the reference, see oracle header), with the networks inside them
  being bit-identical to the reference's (checked by tests/test_oracle.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import drvish_oracle as O  # noqa: E402


def ref_modules_small():
    ref = O.load_reference_modules()
    assert ref is not None, "reference not present"
    out = {}
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        torch.manual_seed(7)
        B, G, L, layers, K, C = 24, 37, 5, [16, 12], 3, 4
        enc = ref.Encoder(G, L, layers).to(dtype)
        dec = ref.NBDecoder(L, G, layers[::-1]).to(dtype)
        # non-trivial BN affine params
        for m in list(enc.modules()) + list(dec.modules()):
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.data.uniform_(0.5, 1.5)
                m.bias.data.uniform_(-0.3, 0.3)
        x = torch.poisson(torch.rand(B, G, dtype=dtype) * 3)
        z_loc, z_scale, l_loc, l_scale = enc(x)
        z = (z_loc + 0.3 * z_scale).detach().requires_grad_(True)
        l = (l_loc + 0.1 * l_scale).detach().requires_grad_(True)
        log_r, logit = dec(z, l)
        gz = torch.autograd.grad((log_r * torch.cos(logit)).sum(), [z, l] + list(dec.parameters()))
        ge = torch.autograd.grad((z_loc * z_scale).sum() + (l_loc * l_scale).sum(), list(enc.parameters()))
        labels = torch.arange(B) % K
        w = torch.randn(L, 1, dtype=dtype)
        b = torch.randn(C, 1, dtype=dtype)
        # logit_mean_sigmoid hard-codes float32 for the scatter target (modules.py:177)
        m = ref.LinearMultiBias.logit_mean_sigmoid(z.detach().float(), w.float(), b.float(), labels)
        out.update({
            f"{tag}.x": x, f"{tag}.z": z.detach(), f"{tag}.l": l.detach(),
            f"{tag}.z_loc": z_loc, f"{tag}.z_scale": z_scale, f"{tag}.l_loc": l_loc, f"{tag}.l_scale": l_scale,
            f"{tag}.log_r": log_r, f"{tag}.logit": logit, f"{tag}.labels": labels, f"{tag}.w": w, f"{tag}.b": b,
            f"{tag}.m": m, f"{tag}.dz": gz[0], f"{tag}.dl": gz[1],
        })
        for (n, _), g in zip(dec.named_parameters(), gz[2:]):
            out[f"{tag}.dec_grad.{n}"] = g
        for (n, _), g in zip(enc.named_parameters(), ge):
            out[f"{tag}.enc_grad.{n}"] = g
        for n, v in enc.state_dict().items():
            out[f"{tag}.enc.{n}"] = v
        for n, v in dec.state_dict().items():
            out[f"{tag}.dec.{n}"] = v
    np.savez_compressed(os.path.join(HERE, "ref_modules_small.npz"),
                        **{k: v.detach().cpu().numpy() for k, v in out.items()})


def oracle_step(name, cfg, seed_data):
    model = O.build_model(cfg, seed=0)
    B = cfg["batch"]
    x = O.synthetic_counts(B, cfg["n_input"], seed=seed_data)
    noise = O.make_noise(model, B, seed=1)
    labels = y = None
    if cfg.get("drugs", 0):
        labels = torch.arange(B) % cfg["classes"]
        g = torch.Generator().manual_seed(99)
        y = [torch.randn(cfg["classes"], cfg["doses"], 1, generator=g) for _ in range(cfg["drugs"])]
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    m64 = O.build_model(cfg, seed=0).double()
    m64.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in sd.items()})
    n64 = {k: ([t.double() for t in v] if isinstance(v, list) else v.double()) for k, v in noise.items()}
    loss64, grads64, terms64 = O.step(m64, x.double(), n64, labels, [t.double() for t in y] if y else None)
    loss32, grads32, _ = O.step(model, x, noise, labels, y)
    out = {"x": x, "loss64": np.float64(loss64), "loss32": np.float64(loss32)}
    for k, v in sd.items():
        out[f"sd.{k}"] = v
    for k, v in noise.items():
        if isinstance(v, list):
            for i, t in enumerate(v):
                out[f"noise.{k}.{i}"] = t
        else:
            out[f"noise.{k}"] = v
    if labels is not None:
        out["labels"] = labels
        for i, t in enumerate(y):
            out[f"y.{i}"] = t
    for k, v in grads64.items():
        out[f"grad64.{k}"] = v
    for k, v in grads32.items():
        out[f"grad32.{k}"] = v
    for k, v in m64.state_dict().items():  # BN running stats after one step
        if "running" in k or "num_batches" in k:
            out[f"after.{k}"] = v
    out["ll64"] = terms64["ll"]
    np.savez_compressed(os.path.join(HERE, name),
                        **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else v) for k, v in out.items()})


def ref_aggmo():
    """ref_aggmo.npz: the reference's OWN optimizer (src/drvish/train/aggmo.py, loaded by file path) run
    for 4 steps on three ragged tensors, each step preceded by the per-parameter
    torch.nn.utils.clip_grad_norm_ that Pyro's PyroOptim applies (clip_args={"clip_norm": c}); the four
    variants cover nesterov on/off, weight decay and clipping that does / does not bite.  Pins SURVEY 8f-1."""
    import importlib.util

    path = "/root/reference/src/drvish/train/aggmo.py"
    spec = importlib.util.spec_from_file_location("ref_aggmo", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    shapes = [(37, 19), (53,), (8, 5, 3)]
    variants = {
        "nesterov_clip": dict(lr=5e-4, betas=[0.0, 0.9, 0.99], nesterov=True, weight_decay=0.0, clip=20.0),
        "plain_clip": dict(lr=1e-2, betas=[0.0, 0.9, 0.99], nesterov=False, weight_decay=0.0, clip=0.5),
        "wd_noclip": dict(lr=3e-3, betas=[0.0, 0.9], nesterov=True, weight_decay=0.01, clip=None),
        "one_beta": dict(lr=1e-3, betas=[0.5], nesterov=False, weight_decay=0.0, clip=100.0),
    }
    out = {}
    for name, v in variants.items():
        g = torch.Generator().manual_seed(99)
        params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
        # one optimizer PER parameter, like PyroOptim
        opts = [mod.AggMo([p], lr=v["lr"], betas=v["betas"], weight_decay=v["weight_decay"], nesterov=v["nesterov"])
                for p in params]
        for i, p in enumerate(params):
            out[f"{name}.p0.{i}"] = p.detach().clone().numpy()
        for step in range(4):
            for i, (p, o) in enumerate(zip(params, opts)):
                scale = 40.0 if step == 1 else 1.0  # step 1: norms well above the clip threshold
                p.grad = torch.randn(p.shape, generator=g) * scale
                out[f"{name}.g{step}.{i}"] = p.grad.clone().numpy()
                if v["clip"] is not None:
                    torch.nn.utils.clip_grad_norm_(p, v["clip"])
                o.step()
                out[f"{name}.p{step + 1}.{i}"] = p.detach().clone().numpy()
    np.savez_compressed(os.path.join(HERE, "ref_aggmo.npz"), **out)


def ref_split_molecules():
    """ref_split_molecules.npz: the reference's OWN MCVDataLoader.split_molecules
    (src/drvish/train/mcv_train.py, loaded by file path) under torch.manual_seed(5).  Pins SURVEY 8f-3's
    loader restatement (oracle.split_molecules)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_mcv_train", "/root/reference/src/drvish/train/mcv_train.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from tests.parity import split_case

    umis, split, comp, overlap = split_case()
    torch.manual_seed(5)
    x0, x1, ls, lc = mod.MCVDataLoader.split_molecules(umis, split, comp, overlap)
    np.savez_compressed(os.path.join(HERE, "ref_split_molecules.npz"), x0=x0.numpy(), x1=x1.numpy(), log_split=ls.numpy(),
                        log_split_complement=lc.numpy())


if __name__ == "__main__":
    ref_modules_small()
    ref_aggmo()
    ref_split_molecules()
    oracle_step("oracle_step_cfg1.npz", O.CONFIGS[1], seed_data=1235)
    oracle_step("oracle_step_tiny.npz",
                dict(n_input=203, n_latent=3, layers=[24, 40], batch=77, drugs=0), seed_data=5)
    print("golden vectors written to", HERE)
