import os, sys, torch
sys.path.insert(0, '.')
from tests import parity as P
import drvish_b200 as D
cfg = dict(n_input=5000, n_latent=10, layers=[256, 256], batch=4096, drugs=8, doses=8, classes=8)
for seed in [int(v) for v in os.environ.get('SEEDS', '1234').split(',')]:
    case = P.make_case(cfg, data_seed=seed)
    loss64, g64, _, _ = P.oracle_step(case, torch.float64)
    for flags in [int(v) for v in os.environ.get('FLAGS', '0').split(',')]:
        model = P.build_cuda_model(case, 'cuda:0')
        elbo = D.FusedTraceELBO(data_parallel=False, extra_flags=flags)
        elbo.set_noise(case['noise'])
        args = [case['x'].cuda(), case['labels'].cuda(), [t.cuda() for t in case['y']]]
        loss = elbo.loss_and_grads(model.model, model.guide, *args)
        errs = {n: P.rel_err(p.grad.double().cpu(), g64[n]) for n, p in model.named_parameters() if float(g64[n].norm()) > 0 and n not in P.zero_grad_biases(2) and not n.startswith('lmb')}
        top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
        print('seed', seed, 'flags', flags, 'loss rel %.2e' % (abs(loss - loss64) / abs(loss64)), [(k, '%.2e' % v) for k, v in top])
