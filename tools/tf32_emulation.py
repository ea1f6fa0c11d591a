"""CPU emulation: which tf32 rounding in the decoder's last layer hurts the gradients?
(fp64 oracle with selected operands of the last Linear rounded to tf32, round-to-nearest)."""
import sys, torch
sys.path.insert(0, '.')
from tests import parity as P
from oracle import drvish_oracle as O

def rna(t):
    f = t.float().contiguous()
    i = f.view(torch.int32)
    r = ((i + 0x1000) & ~0x1FFF)
    return r.view(torch.float32).to(t.dtype)

class LastLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, act, W, b, mode):
        ctx.mode = mode
        ctx.save_for_backward(act, W)
        a, w = (rna(act), rna(W)) if 'F' in mode else (act, W)
        return a @ w.t() + b
    @staticmethod
    def backward(ctx, g):
        act, W = ctx.saved_tensors
        mode = ctx.mode
        gW = (rna(g).t() @ rna(act)) if 'W' in mode else g.t() @ act
        gA = (rna(g) @ rna(W)) if 'A' in mode else g @ W
        return gA, gW, g.sum(0), None

cfg = dict(n_input=5000, n_latent=10, layers=[256, 256], batch=4096, drugs=8, doses=8, classes=8)
case = P.make_case(cfg, data_seed=int(sys.argv[1]) if len(sys.argv) > 1 else 1234)
loss64, g64, _, _ = P.oracle_step(case, torch.float64)
for mode in ('F', 'W', 'A', 'FWA'):
    m = O.build_model(cfg, seed=0).double()
    m.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in case['model'].state_dict().items()})
    lin = m.decoder.fc[8]
    class Wrap(torch.nn.Module):
        def __init__(s, lin): super().__init__(); s.weight, s.bias = lin.weight, lin.bias
        def forward(s, x): return LastLinear.apply(x, s.weight, s.bias, mode)
    m.decoder.fc[8] = Wrap(lin)
    noise = {k: ([t.double() for t in v] if isinstance(v, list) else v.double()) for k, v in case['noise'].items()}
    loss, grads, _ = O.step(m, case['x'].double(), noise, case['labels'], [t.double() for t in case['y']])
    errs = {n: P.rel_err(g, g64[n]) for n, g in grads.items() if n in g64 and float(g64[n].norm()) > 0 and n not in P.zero_grad_biases(2) and not n.startswith('lmb')}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(mode, 'loss rel %.2e' % (abs(loss - loss64) / abs(loss64)), [(k, '%.2e' % v) for k, v in top], flush=True)
