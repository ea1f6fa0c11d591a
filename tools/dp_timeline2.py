"""Data-parallel step timeline on rank 0 (CUPTI via torch.profiler): every kernel of one graph replay (between two
k_draw_step_noise launches) that runs longer than MIN us or belongs to the collective.
usage: torchrun --nproc-per-node N tools/dp_timeline2.py   (env CFG, MIN)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import drvish_b200 as D

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(bench.CONFIGS[int(os.environ.get("CFG", "3"))])
model = bench.build_native_model(cfg, dev)
rt = model._runtime()
elbo = D.FusedTraceELBO(seed=1, data_parallel=True, use_cuda_graph=True)
xs, labels, y = bench.synth_batches(cfg, 4, 7 + rank, dev)
yflat = elbo._flatten_y(rt, y, dev) if cfg["drugs"] else None
rt.ensure_flat()
elbo.dp_prepare(rt)


def step(i):
    elbo.dp_step(rt, xs[i % 4], labels, yflat)
    elbo.step_index += 1


for i in range(16):
    step(i)
torch.cuda.synchronize()
dist.barrier()
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(6):
        step(i)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    starts = [i for i, e in enumerate(evs) if "k_draw_step_noise" in e.name]
    seg = evs[starts[3]:starts[4]]
    t0 = seg[0].time_range.start
    mn = float(os.environ.get("MIN", "12"))
    for e in seg:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        if d > mn or "allreduce" in e.name.lower():
            print("%8.1f %8.1f %8.1f  %s" % (s, d, s + d, e.name.replace("drv::(anonymous namespace)::", "").replace("void ", "")[:80]))
    print("step span us", max(e.time_range.end for e in seg) - t0, " next step starts at", evs[starts[4]].time_range.start - t0)
dist.destroy_process_group()
