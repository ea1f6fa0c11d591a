"""Where does the gradient error of a dispersion-sweep case enter?  Runs one parity case under the debug flags that
switch single tensor-core stages back to the fp32 CUDA-core kernels (DRVISH_B200_DEBUG_FLAGS: 4 trunk backward,
8 trunk forward, 16 decoder, 32 encoder input layer) and under the decoder's operand switches.
    python tools/disp_probe.py [case] [shift]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P  # noqa: E402
from tests.test_cuda_parity import CASES  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "mid"
shift = float(sys.argv[2]) if len(sys.argv) > 2 else -8.0
case = P.make_case(CASES[name], data_seed=11, log_r_shift=shift)
runs = [("default", {}), ("trunk bwd fp32 (4)", {"DRVISH_B200_DEBUG_FLAGS": "4"}), ("trunk fwd fp32 (8)", {"DRVISH_B200_DEBUG_FLAGS": "8"}),
        ("trunk fp32 (12)", {"DRVISH_B200_DEBUG_FLAGS": "12"}), ("decoder fp32 (16)", {"DRVISH_B200_DEBUG_FLAGS": "16"}),
        ("encoder-in fp32 (32)", {"DRVISH_B200_DEBUG_FLAGS": "32"}), ("decoder tf32 operands", {"DRV_DEC_TF32": "1"}),
        ("unfused dW", {"DRV_DEC_FUSED_DW": "0"}), ("all fp32", {"DRVISH_B200_DEBUG_FLAGS": "1"})]
for label, env in runs:
    for k in ("DRVISH_B200_DEBUG_FLAGS", "DRV_DEC_TF32", "DRV_DEC_FUSED_DW"):
        os.environ.pop(k, None)
    os.environ.update(env)
    rep = P.compare_step(case)
    g = rep["grads"]
    print("%-24s path=%d loss_rel=%.1e worst=%.2e (%s)" % (label, rep["path"], rep["loss_rel"], rep["worst_grad_rel"], rep["worst_grad_name"]))
    print("    " + " ".join("%s=%.1e" % (k.replace("decoder.fc.", "d").replace("encoder.fc.", "e").replace("weight", "w").replace("bias", "b"), v[1])
                            for k, v in g.items() if v[0] == "rel"))
