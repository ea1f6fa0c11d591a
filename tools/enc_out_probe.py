"""Encoder forward outputs (z_loc, z_scale, l_loc, l_scale) of the CUDA path vs the fp64 oracle, tensor-core input layer
on / off (DRVISH_B200_DEBUG_FLAGS=32), and the decoder-trunk gradient error over a log_r shift sweep.
    python tools/enc_out_probe.py [case]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P  # noqa: E402
from tests.test_cuda_parity import CASES  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mid"
case = P.make_case(CASES[name], data_seed=11)
_, _, _, m64 = P.oracle_step(case, torch.float64)
with torch.no_grad():
    ref = m64.encoder(case["x"].double())
ref32 = None
for flags in ("0", "32"):
    os.environ["DRVISH_B200_DEBUG_FLAGS"] = flags
    model = P.build_cuda_model(case, "cuda:0")
    out = model.encoder(case["x"].cuda())
    print("flags", flags, " ".join("%s: rel %.2e max|d| %.2e" % (n, P.rel_err(o.cpu(), r), float((o.cpu().double() - r).abs().max()))
                                    for n, o, r in zip(("z_loc", "z_scale", "l_loc", "l_scale"), out, ref)))
for shift in (0.0, -4.0, -6.0, -8.0):
    c = P.make_case(CASES[name], data_seed=11, log_r_shift=shift)
    for flags in ("0", "32"):
        os.environ["DRVISH_B200_DEBUG_FLAGS"] = flags
        rep = P.compare_step(c)
        g = rep["grads"]
        print("shift %5.1f flags %2s worst %.2e | " % (shift, flags, rep["worst_grad_rel"]) +
              " ".join("%s=%.1e" % (k.replace("decoder.fc.", "d").replace("weight", "w").replace("bias", "b"), v[1])
                       for k, v in g.items() if v[0] == "rel" and k.startswith("decoder")))
