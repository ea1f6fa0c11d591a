import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P
from tests.test_cuda_parity import CASES
name = sys.argv[1] if len(sys.argv) > 1 else "mid"
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 11
case = P.make_case(CASES[name], data_seed=seed)
for flags in ("0", "0", "64", "512", "1024", "2048", "32", "96"):
    os.environ["DRVISH_B200_DEBUG_FLAGS"] = flags
    rep = P.compare_step(case)
    g = rep["grads"]
    print("flags %4s worst %.2e | " % (flags, rep["worst_grad_rel"]) +
          " ".join("%s=%.1e" % (k.replace("decoder.fc.", "d").replace("encoder.fc.", "e").replace("weight", "w").replace("bias", "b"), v[1])
                   for k, v in g.items() if v[0] == "rel"))
