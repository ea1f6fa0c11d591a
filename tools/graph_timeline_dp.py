"""Timeline (CUPTI via torch.profiler; not a bench number) of data-parallel steps on rank 0: one CUDA-graph replay
per step with the peer-memory all-reduce as its last node.  Prints every kernel of the third profiled step that is
longer than 8 us or belongs to the collective, and the gaps around the all-reduce.
usage: torchrun --nproc-per-node N tools/graph_timeline_dp.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
import drvish_b200 as D

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
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
    ar = [i for i, e in enumerate(evs) if "allreduce" in e.name.lower()]
    print("all-reduce kernels:", len(ar))
    for k in range(1, len(ar)):
        a, b = ar[k - 1], ar[k]
        e = evs[b]
        prev_end = max(x.time_range.end for x in evs[a + 1:b])
        first_next = evs[a + 1].time_range.start
        print("step %d: span %.1f us (AR end to AR end); AR duration %.1f; gap before AR %.1f; gap after previous AR %.1f" % (
            k, e.time_range.end - evs[a].time_range.end, e.time_range.end - e.time_range.start,
            e.time_range.start - prev_end, first_next - evs[a].time_range.end))
    if len(ar) >= 4:
        seg = evs[ar[2] + 1:ar[3] + 1]
        t0 = seg[0].time_range.start
        for e in seg:
            s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
            if d > 8 or "allreduce" in e.name.lower() or "Memset" in e.name or "Memcpy" in e.name:
                print("%8.1f %8.1f %8.1f  %s" % (s, d, s + d, e.name[:90]))
dist.destroy_process_group()
