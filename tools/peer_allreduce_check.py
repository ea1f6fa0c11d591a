"""Correctness + timing of drv_peer_allreduce against NCCL.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/peer_allreduce_check.py [n_floats]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from drvish_b200.parallel import PeerAllReduce

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
world = dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_002_698
peer = PeerAllReduce.create(n, dev)
assert peer is not None, "peer all-reduce unavailable"
g = torch.Generator(device=dev).manual_seed(100 + rank)
worst = 0.0
for it in range(20):
    peer.data.copy_(torch.randn(n, device=dev, generator=g) * (1.0 + it))
    ref = peer.data.clone()
    dist.all_reduce(ref)
    peer.all_reduce()
    torch.cuda.synchronize()
    err = float((peer.data - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    # bit-identical on every rank
    chk = peer.data.view(torch.int32).sum(dtype=torch.int64).reshape(1)
    all_chk = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(all_chk, chk)
    assert len({int(c) for c in all_chk}) == 1, "ranks disagree"
assert worst < 1e-6, worst


def timeit(fn, iters=50):
    for _ in range(5):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t) * 1e3


peer.data.zero_()
t_peer = timeit(peer.all_reduce)
buf = torch.zeros(n, device=dev)
t_nccl = timeit(lambda: dist.all_reduce(buf))
if rank == 0:
    print(f"world {world} n_floats {n} ({4 * n / 1e6:.1f} MB): worst rel err vs NCCL {worst:.1e}, bit-identical across ranks; "
          f"peer kernel {t_peer:.1f} us, NCCL {t_nccl:.1f} us")
dist.destroy_process_group()
