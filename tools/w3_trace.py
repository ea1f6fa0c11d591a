"""Timeline of k_dec3w's CTA 0 (fused phase 3 + dW kernel), from globaltimer stamps (DRV_W3_TRACE=1):
    python tools/w3_trace.py [config]
columns per tile (us since the first stamp): producer act issue / ds' issue, MMA1 issue, MMA2 issue, store start / read done,
epilogue: ds' tile seen, MMA1 seen, TMEM loaded, tile done."""
import ctypes as C
import os
import sys

os.environ["DRV_W3_TRACE"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import drvish_b200 as D  # noqa: E402
from drvish_b200 import _lib  # noqa: E402

cfg = dict(bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 3])
dev = torch.device("cuda", 0)
model = bench.build_native_model(cfg, dev)
xs, labels, y = bench.synth_batches(cfg, 2, 1234, dev)
elbo = D.FusedTraceELBO(seed=1, data_parallel=False, use_cuda_graph=False)
for i in range(4):
    args = [xs[i % 2]] + ([labels, y] if cfg["drugs"] else [])
    elbo.loss_and_grads(model.model, model.guide, *args)
torch.cuda.synchronize()
lib = _lib.load()
T, E = 24, 10
buf = (C.c_ulonglong * (T * E))()
n = lib.drv_debug_w3_trace(buf, T * E)
v = [buf[i] for i in range(n)]
t0 = min(x for x in v if x)
names = ["act_iss", "ds_iss", "mma1", "mma2", "st_beg", "st_rd", "e_ds", "e_mma1", "e_tmem", "e_done"]
print("tile " + " ".join("%8s" % s for s in names))
for slot, name in ((21, "first CTA"), (22, "middle CTA"), (23, "last CTA")):
    row = v[slot * E:slot * E + 5]
    print(name, "entry / pdl wait done / set-up done / walk done / exit:", " ".join("%8.2f" % ((x - t0) / 1e3) if x else "-" for x in row))
for it in range(21):
    row = v[it * E:(it + 1) * E]
    if not any(row):
        break
    print("%4d " % it + " ".join("%8.2f" % ((x - t0) / 1e3) if x else "       -" for x in row))
