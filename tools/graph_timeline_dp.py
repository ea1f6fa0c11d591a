"""Timeline of one data-parallel step (two-part graph replay + NCCL all-reduces) on rank 0.
usage: torchrun --nproc-per-node 2 tools/graph_timeline_dp.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
import drvish_b200 as D

rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = dict(bench.CONFIGS[3])
model = bench.build_native_model(cfg, dev)
rt = model._runtime()
elbo = D.FusedTraceELBO(seed=1, data_parallel=True, use_cuda_graph=True)
xs, labels, y = bench.synth_batches(cfg, 4, 7 + rank, dev)
yflat = elbo._flatten_y(rt, y, dev)
rt.ensure_flat()

def step(i):
    xb = xs[i % 4]
    elbo._enqueue(rt, xb, labels, yflat, True, part=1)
    w1 = dist.all_reduce(rt.flat_grads[rt.decoder_offset:], op=dist.ReduceOp.SUM, async_op=True)
    elbo._enqueue(rt, xb, labels, yflat, True, part=2)
    w2 = dist.all_reduce(rt.flat_grads[:rt.decoder_offset], op=dist.ReduceOp.SUM, async_op=True)
    w1.wait(); w2.wait()
    elbo.step_index += 1

for i in range(16): step(i)
torch.cuda.synchronize(); dist.barrier()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(4): step(i)
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    names = [e.name for e in evs]
    first = names[0]
    starts = [i for i, n in enumerate(names) if n == first]
    seg = evs[starts[2]:starts[3]] if len(starts) > 3 else evs[starts[-1]:]
    t0 = seg[0].time_range.start
    for e in seg:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        if d > 8 or "nccl" in e.name.lower() or "Memcpy" in e.name:
            print("%8.1f %8.1f %8.1f  %s" % (s, d, s + d, e.name[:100]))
    print("span", max(e.time_range.end for e in seg) - t0)
dist.destroy_process_group()
