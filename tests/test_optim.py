"""Fused AggMo optimizer.  CPU: the oracle restatement against golden vectors
produced by the reference's own aggmo.py (tests/golden/make_golden.py:ref_aggmo).  GPU: drv_aggmo_step
(through FusedAggMo) against the oracle on a real model layout.  Tolerance: fp32 elementwise
arithmetic with fused multiply-adds -> 2e-6 relative to the tensor's max magnitude."""
import os

import numpy as np
import pytest
import torch

from oracle.aggmo_oracle import AggMoOracle

VARIANTS = {
    "nesterov_clip": dict(lr=5e-4, betas=[0.0, 0.9, 0.99], nesterov=True, weight_decay=0.0, clip=20.0),
    "plain_clip": dict(lr=1e-2, betas=[0.0, 0.9, 0.99], nesterov=False, weight_decay=0.0, clip=0.5),
    "wd_noclip": dict(lr=3e-3, betas=[0.0, 0.9], nesterov=True, weight_decay=0.01, clip=None),
    "one_beta": dict(lr=1e-3, betas=[0.5], nesterov=False, weight_decay=0.0, clip=100.0),
}


@pytest.mark.parametrize("name", list(VARIANTS))
def test_oracle_matches_reference_aggmo_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "ref_aggmo.npz"))
    v = VARIANTS[name]
    params = [torch.from_numpy(g[f"{name}.p0.{i}"]).clone() for i in range(3)]
    opt = AggMoOracle(params, v["lr"], v["betas"], v["weight_decay"], v["nesterov"], v["clip"])
    for step in range(4):
        opt.step([torch.from_numpy(g[f"{name}.g{step}.{i}"]) for i in range(3)])
        for i, p in enumerate(params):
            torch.testing.assert_close(p, torch.from_numpy(g[f"{name}.p{step + 1}.{i}"]), rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
@pytest.mark.parametrize("drugs", [0, 3])
def test_fused_aggmo_matches_oracle(name, drugs):
    import drvish_b200 as D

    v = VARIANTS[name]
    torch.manual_seed(3)
    kw = dict(n_input=203, n_latent=5, layers=[24, 40])
    if drugs:
        model = D.DRNBVAE(n_classes=3, n_drugs=drugs, n_conditions=[4, 2, 5], **kw).cuda()
    else:
        model = D.NBVAE(**kw).cuda()
    rt = model._runtime()
    rt.ensure_flat()
    params = list(rt._params)
    ref = [p.detach().cpu().clone() for p in params]
    oracle = AggMoOracle(ref, v["lr"], v["betas"], v["weight_decay"], v["nesterov"], v["clip"])
    opt = D.FusedAggMo(model, lr=v["lr"], betas=v["betas"], weight_decay=v["weight_decay"], nesterov=v["nesterov"],
                       clip_norm=v["clip"])
    g = torch.Generator().manual_seed(17)
    for step in range(4):
        grads = [torch.randn(p.shape, generator=g) * (40.0 if step == 1 else 1.0) for p in params]
        rt.flat_grads.zero_()
        for view, gr in zip(rt._grad_views, grads):
            view.copy_(gr.cuda())  # views of the flat gradient buffer (what param.grad points at)
        opt.step()
        oracle.step(grads)
        for p, r in zip(params, ref):
            err = float((p.detach().cpu() - r).abs().max()) / max(float(r.abs().max()), 1e-12)
            assert err <= 2e-6, (step, tuple(p.shape), err)


@pytest.mark.gpu
def test_fused_aggmo_is_a_torch_optimizer_with_schedulers_and_state():
    import drvish_b200 as D

    model = D.NBVAE(n_input=64, n_latent=4, layers=[32]).cuda()
    opt = D.FusedAggMo.from_exp_form(model, lr=3e-3, a=0.1, k=3, nesterov=True, clip_norm=20.0)
    assert opt.param_groups[0]["betas"] == [0.0, 0.9, 0.99] and abs(opt.param_groups[0]["lr"] - 1e-3) < 1e-12
    sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=4)
    rt = model._runtime()
    rt.ensure_flat()
    before = rt.flat_params.clone()
    rt.flat_grads.normal_()
    lrs = []
    for _ in range(5):
        opt.step()
        sched.step()
        lrs.append(opt.param_groups[0]["lr"])
    assert lrs[0] < 1e-3 and lrs[3] == pytest.approx(1e-3)  # cosine decay, restart after T_0 epochs
    assert not torch.equal(before, rt.flat_params)
    sd = opt.state_dict()
    opt2 = D.FusedAggMo(model, lr=3e-3, nesterov=True, clip_norm=20.0)
    opt2.load_state_dict(sd)
    assert torch.equal(opt2._momentum, opt._momentum)
    # the Pyro-facing callable
    po = D.PyroFusedAggMo(model, {"lr": 5e-4, "betas": [0.0, 0.9, 0.99], "nesterov": True}, {"clip_norm": 20.0})
    po(list(model.parameters()))
