"""A few direct (no CUDA graph) ELBO steps of one bench config, for ncu / compute-sanitizer:
    python tools/prof_step.py [config] [steps] [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import drvish_b200 as D  # noqa: E402

cfg = dict(bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 3])
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if len(sys.argv) > 3:
    cfg["batch"] = int(sys.argv[3])
dev = torch.device("cuda", 0)
model = bench.build_native_model(cfg, dev)
xs, labels, y = bench.synth_batches(cfg, 2, 1234, dev)
elbo = D.FusedTraceELBO(seed=1, data_parallel=False, use_cuda_graph=False)
for i in range(steps):
    args = [xs[i % 2]] + ([labels, y] if cfg["drugs"] else [])
    loss = elbo.loss_and_grads(model.model, model.guide, *args)
torch.cuda.synchronize()
print("loss", loss, "launches/step", model._runtime().last_launches, "path", model._runtime().last_path)
