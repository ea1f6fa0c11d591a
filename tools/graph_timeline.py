"""Timeline of ONE CUDA-graph replay of the step (kernel start / duration / stream) through
torch.profiler (CUPTI), to see the real critical path with the side stream and PDL active.
usage: python tools/graph_timeline.py [config]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import drvish_b200 as D

cfg = dict(bench.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 3])
dev = torch.device("cuda", 0)
model = bench.build_native_model(cfg, dev)
rt = model._runtime()
elbo = D.FusedTraceELBO(seed=1, data_parallel=False, use_cuda_graph=True)
xs, labels, y = bench.synth_batches(cfg, 4, 7, dev)
yflat = elbo._flatten_y(rt, y, dev) if cfg["drugs"] else None
rt.ensure_flat()
for i in range(12):
    elbo._enqueue(rt, xs[i % 4], labels, yflat, True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(4):
        elbo._enqueue(rt, xs[i % 4], labels, yflat, True)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into replays by the first kernel name
names = [e.name for e in evs]
first = names[0]
starts = [i for i, n in enumerate(names) if n == first]
seg = evs[starts[2]:starts[3]] if len(starts) > 3 else evs[starts[-1]:]
t0 = seg[0].time_range.start
end = 0
print("%8s %8s %8s  %s" % ("start", "dur", "end", "kernel"))
for e in seg:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    gap = s - end
    end = max(end, s + d)
    print("%8.1f %8.1f %8.1f  %s%s" % (s, d, s + d, ("[gap %.1f] " % gap) if gap > 1.0 else "", e.name[:90]))
print("replay span us", end)
