"""Host packer throughput on this box: python tools/pack_bench.py  (pinned source and destination, like the prefetcher)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drvish_b200.data import PackedCounts8, PackedCounts
from drvish_b200 import _lib
print("isa", _lib.load().drv_host_pack_isa().decode(), "cpus", os.cpu_count())
g = torch.Generator().manual_seed(0)
xs = [torch.poisson(torch.full((4096, 5000), 1.5), generator=g).pin_memory() for _ in range(3)]
n = xs[0].numel()
bufs = [PackedCounts8.buffers(n) for _ in range(3)]
for nt in (4, 8, 12, 14, 15, 16):
    for k in range(3):
        PackedCounts8.pack(xs[k], n_threads=nt, out=bufs[k])
    t = time.perf_counter()
    for it in range(12):
        PackedCounts8.pack(xs[it % 3], n_threads=nt, out=bufs[it % 3])
    dt = (time.perf_counter() - t) / 12
    print(f"u8  {nt:2d} threads {dt*1e3:6.2f} ms/tile  {4*n/dt/1e9:6.1f} GB/s read")
