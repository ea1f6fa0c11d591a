import os, sys, torch
sys.path.insert(0, '.')
from tests import parity as P
cfg = dict(n_input=5000, n_latent=10, layers=[256, 256], batch=4096, drugs=8, doses=8, classes=8)
case = P.make_case(cfg, data_seed=1234)
_, _, _, m64 = P.oracle_step(case, torch.float64)
m64 = P.O.build_model(cfg, seed=0).double()
m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in case['model'].state_dict().items()})
with torch.no_grad():
    ref = m64.encoder(case['x'].double())
    h1 = m64.encoder.fc[:3](torch.sqrt(case['x'].double()))
    m32 = P.O.build_model(cfg, seed=0); m32.load_state_dict(case['model'].state_dict())
    o32 = m32.encoder(case['x'])
for flags in ('0', '32', '1'):
    os.environ['DRVISH_B200_DEBUG_FLAGS'] = flags
    model = P.build_cuda_model(case, 'cuda:0')
    with torch.no_grad():
        out = model.encoder(case['x'].cuda())
    print('flags', flags, ['%.2e' % P.rel_err(o.cpu(), r) for o, r in zip(out, ref)], 'fp32 oracle:', ['%.2e' % P.rel_err(o, r) for o, r in zip(o32, ref)])
