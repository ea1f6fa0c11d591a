"""CPU emulation (numpy float32) of common.cuh:nb_elems against fp64 scipy, over dispersion regimes.
Run here (no GPU): python tools/nb_math_probe.py.  Used to pick the formulas / break-even of the
large-r route before spending GPU time."""
import numpy as np
from scipy.special import gammaln, digamma

f = np.float32
LN2 = f(0.6931471805599453)
LOG2E = f(1.4426950408889634)


def lg2(v):
    return np.log2(v.astype(np.float64)).astype(np.float32)


def ex2(v):
    return np.exp2(v.astype(np.float64)).astype(np.float32)


def rcp(v):
    return (1.0 / v.astype(np.float64)).astype(np.float32)


def kernel_r01(lsl, rho, x, eps):
    """round-1 nb_elems (Stirling differences at shift 4)"""
    rexp = ex2(rho * LOG2E)
    r = rexp + f(eps)
    ell = lsl - rho
    e = ex2(-np.abs(ell) * LOG2E)
    t = f(1) + e
    sp = lg2(t) * LN2 + np.maximum(ell, f(0))
    rx = r + x
    y0, y1 = r + f(4), rx + f(4)
    u0 = r * (r + f(3)); w0 = u0 + f(2); q0 = u0 * w0
    u1 = rx * (rx + f(3)); w1 = u1 + f(2); q1 = u1 * w1
    i0, i1 = rcp(y0), rcp(y1)
    l0, l1 = lg2(y0) * LN2, lg2(y1) * LN2
    i0s, i1s = i0 * i0, i1 * i1
    iq0 = rcp(q0)
    c0 = i0 * (i0s * f(-2.7777777778e-3) + f(8.3333333333e-2))
    c1 = i1 * (i1s * f(-2.7777777778e-3) + f(8.3333333333e-2))
    lgd = ((y1 - f(0.5)) * l1 + (-(y0 - f(0.5)) * l0)) - x + (c1 - c0) - lg2(q1 * iq0) * LN2
    iq1 = rcp(q1)
    g0 = l0 - f(0.5) * i0 - i0s * (i0s * f(-8.3333333333e-3) + f(8.3333333333e-2))
    g1 = l1 - f(0.5) * i1 - i1s * (i1s * f(-8.3333333333e-3) + f(8.3333333333e-2))
    qd0 = (u0 + w0) * (f(2) * r + f(3))
    qd1 = (u1 + w1) * (f(2) * rx + f(3))
    psd = (g1 - g0) - (qd1 * iq1 - qd0 * iq0)
    z = x == 0
    lgd = np.where(z, f(0), lgd)
    psd = np.where(z, f(0), psd)
    lgf = gammaln(x.astype(np.float64) + 1).astype(np.float32)
    cnt = lgd - lgf
    ll = (x * ell + cnt) - rx * sp
    inv = rcp(t)
    sig = np.where(ell >= 0, inv, e * inv)
    a = x - rx * sig
    drho = rexp * (psd - sp) - a
    return ll, a, drho


def exact(lsl, rho, x, eps):
    lsl, rho, x = lsl.astype(np.float64), rho.astype(np.float64), x.astype(np.float64)
    r = np.exp(rho) + eps
    ell = lsl - rho
    sp = np.logaddexp(0, ell)
    ll = x * ell - (x + r) * sp + gammaln(r + x) - gammaln(r) - gammaln(1 + x)
    sig = 1 / (1 + np.exp(-ell))
    a = x - (x + r) * sig
    drho = np.exp(rho) * (-sp + digamma(r + x) - digamma(r)) - a
    return ll, a, drho


def torch_fp32(lsl, rho, x, eps):
    import torch
    from torch.distributions import NegativeBinomial
    rho_t = torch.tensor(rho, requires_grad=True)
    lsl_t = torch.tensor(lsl, requires_grad=True)
    xt = torch.tensor(x)
    nb = NegativeBinomial(total_count=torch.exp(rho_t) + eps, logits=lsl_t - rho_t)
    ll = nb.log_prob(xt)
    ll.sum().backward()
    return ll.detach().numpy(), lsl_t.grad.numpy(), rho_t.grad.numpy()


def report(fn, name, shifts=(-8, -4, 0, 4, 8), n=200000, seed=0):
    rng = np.random.default_rng(seed)
    print(f"--- {name}")
    for s in shifts:
        rho = (rng.normal(0, 0.5, n) + s).astype(np.float32)
        mu = np.exp(rng.normal(0, 1.2, n))                      # NB mean ~ 1
        x = rng.poisson(mu * rng.gamma(2.0, 0.5, n)).astype(np.float32)
        lsl = (np.log(mu) + 0 * rho).astype(np.float32)         # ell = log(mu) - rho  (mean = r e^ell = mu)
        ll, a, dr = fn(lsl, rho, x, 1e-6)
        ll64, a64, dr64 = exact(lsl, rho, x, 1e-6)
        rel = lambda v, w: np.linalg.norm(v.astype(np.float64) - w) / np.linalg.norm(w)
        print(f"shift {s:+d}: r~{np.exp(s):.1e}  ll sum rel {abs(ll.astype(np.float64).sum() - ll64.sum()) / abs(ll64.sum()):.2e}"
              f"  ll elem {rel(ll, ll64):.2e}  a {rel(a, a64):.2e}  drho {rel(dr, dr64):.2e}"
              f"  drho colsum rel {abs(dr.astype(np.float64).sum() - dr64.sum()) / abs(dr64.sum()):.2e}")


if __name__ == "__main__":
    report(torch_fp32, "torch fp32 NegativeBinomial.log_prob + autograd (the reference's own arithmetic)")
    report(kernel_r01, "round-1 kernel formulas (fp32 emulation)")


def kernel_r02(lsl, rho, x, eps, merge_t=True):
    """round-2 nb_pair: differences formed analytically (exactly 0 at x = 0), merged reciprocals,
    a = x sigmoid(-ell) - r sigmoid(ell), ll = cnt - x softplus(-ell) - r softplus(ell)."""
    rexp = ex2(rho * LOG2E)
    r = rexp + f(eps)
    nr = -r
    ell = lsl - rho
    e = ex2(-np.abs(ell) * LOG2E)
    t = f(1) + e
    spp = lg2(t) * LN2 + np.maximum(ell, f(0))
    nspn = ell - spp
    rx = r + x
    y0, y1 = r + f(4), rx + f(4)
    u0 = r * (r + f(3)); w0 = u0 + f(2); q0 = u0 * w0
    u1 = rx * (rx + f(3)); w1 = u1 + f(2); q1 = u1 * w1
    P = y0 * y1
    if merge_t:
        rP = rcp(P * t)
        i01 = rP * t
        inv = rP * P
    else:
        i01 = rcp(P)
        inv = rcp(t)
    Q = q0 * q1
    rQ = rcp(Q)
    ratio = (q1 * rQ) * q1
    l0, l1 = lg2(y0), lg2(y1)
    lgr = lg2(ratio)
    dl = l1 - l0
    main = x * l1 + (y0 - f(0.5)) * dl
    dpos = x * i01                      # 1/y0 - 1/y1
    s = (y0 + y1) * i01                 # 1/y0 + 1/y1
    S2 = s * s
    A1 = S2 - i01                       # i0^2 + i0 i1 + i1^2
    A2 = S2 - f(2) * i01                # i0^2 + i1^2
    cc = dpos * (A1 * f(1 / 360) + f(-1 / 12))
    lgf = gammaln(x.astype(np.float64) + 1).astype(np.float32)
    cnt = ((main - lgr) * LN2 + cc) - (x + lgf)
    gdi = dpos * (s * (A2 * f(-1 / 120) + f(1 / 12)) + f(0.5))
    gd = dl * LN2 + gdi
    qd0n = (u0 + f(1)) * (r * f(-2) + f(-3))
    qd1 = (u1 + f(1)) * (rx * f(2) + f(3))
    nn = qd1 * q0 + qd0n * q1
    zq = nn * rQ
    psd = zq * f(-2) + gd
    ll = x * nspn + (nr * spp + cnt)
    einv = e * inv
    pos = ell >= 0
    sg = np.where(pos, inv, einv)
    sn = np.where(pos, einv, inv)
    a = x * sn + nr * sg
    ndrho = rexp * (spp - psd) + a
    return ll, a, -ndrho


if __name__ == "__main__":
    report(kernel_r02, "round-2 kernel formulas (fp32 emulation)")
    report(lambda *a: kernel_r02(*a, merge_t=False), "round-2, rcp(t) kept separate")
    print("--- fast-path error against r (x ~ Poisson, mean 2): where does the generic route have to take over?")
    rng = np.random.default_rng(1)
    n = 100000
    for lr in (5, 7, 8, 9, 9.7, 10.3):
        rho = np.full(n, lr, np.float32)
        mu = np.exp(rng.normal(0.5, 0.7, n))
        x = rng.poisson(mu).astype(np.float32)
        lsl = np.log(mu).astype(np.float32)
        ll, a, dr = kernel_r02(lsl, rho, x, 1e-6)
        ll64, a64, dr64 = exact(lsl, rho, x, 1e-6)
        t32 = torch_fp32(lsl, rho, x, 1e-6)
        print(f"r={np.exp(lr):9.0f}: ll abs err mean {np.abs(ll - ll64).mean():.2e} (torch fp32 {np.abs(t32[0] - ll64).mean():.2e}), "
              f"ll sum rel {abs(ll.astype(np.float64).sum() - ll64.sum()) / abs(ll64.sum()):.2e} (torch {abs(t32[0].astype(np.float64).sum() - ll64.sum()) / abs(ll64.sum()):.2e}), "
              f"a rel {np.linalg.norm(a - a64) / np.linalg.norm(a64):.2e}, drho abs err mean {np.abs(dr - dr64).mean():.2e} (|drho| mean {np.abs(dr64).mean():.2e})")


def generic_large_r(lsl, rho, x, eps):
    """nb_elem_generic, r >= 32 branch (accurate libdevice functions modelled by float64 -> float32 rounding)."""
    c = lambda v: np.asarray(v, np.float64).astype(np.float32)
    rexp = c(np.exp(rho.astype(np.float64)))
    r = rexp + f(eps)
    ell = lsl - rho
    e = c(np.exp(-np.abs(ell.astype(np.float64))))
    L = c(np.log1p(e.astype(np.float64)))
    spp, spn = L + np.maximum(ell, f(0)), L + np.maximum(-ell, f(0))
    rx = r + x
    i0, i1 = f(1) / r, f(1) / rx
    l1p = c(np.log1p((x * i0).astype(np.float64)))
    d = -x * i0 * i1
    s, p = i0 + i1, i0 * i1
    s2 = s * s
    cnt = (r - f(0.5)) * l1p + x * (c(np.log(rx.astype(np.float64))) - f(1)) + d * (f(1 / 12) - (s2 - p) * f(1 / 360))
    psd = l1p - d * (f(0.5) + s * (f(1 / 12) - (s2 - f(2) * p) * f(1 / 120)))
    cnt = np.where(x == 0, f(0), cnt - gammaln(x.astype(np.float64) + 1).astype(np.float32))
    psd = np.where(x == 0, f(0), psd)
    ll = cnt - x * spn - r * spp
    inv = f(1) / (f(1) + e)
    sg = np.where(ell >= 0, inv, e * inv)
    sn = np.where(ell >= 0, e * inv, inv)
    a = x * sn - r * sg
    return ll, a, -(rexp * (spp - psd) + a)


if __name__ == "__main__":
    print("--- generic route, r >= 32 branch")
    rng = np.random.default_rng(1)
    n = 100000
    for lr in (3.6, 5, 8, 11, 14):
        rho = (np.full(n, lr) + rng.normal(0, 0.3, n)).astype(np.float32)
        mu = np.exp(rng.normal(0.5, 0.7, n))
        x = rng.poisson(mu).astype(np.float32)
        lsl = np.log(mu).astype(np.float32)
        ll, a, dr = generic_large_r(lsl, rho, x, 1e-6)
        ll64, a64, dr64 = exact(lsl, rho, x, 1e-6)
        print(f"r~{np.exp(lr):9.0f}: ll abs err mean {np.abs(ll - ll64).mean():.2e}, ll sum rel {abs(ll.astype(np.float64).sum() - ll64.sum()) / abs(ll64.sum()):.2e}, "
              f"a rel {np.linalg.norm(a - a64) / np.linalg.norm(a64):.2e}, drho rel {np.linalg.norm(dr - dr64) / np.linalg.norm(dr64):.2e}")
