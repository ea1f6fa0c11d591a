"""CPU tests of the host side: drop-in module surface (constructor keywords, attributes,
state_dict keys as in reference nbvae.py / drnbvae.py / modules.py), loud failure without CUDA,
and the data-parallel reduction under a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import drvish_b200 as D
from oracle import drvish_oracle as O


def test_constructor_surface_and_state_dict_keys_match_reference():
    m = D.DRNBVAE(n_input=40, n_classes=3, n_drugs=2, n_conditions=[4, 5], n_latent=6, layers=[16, 8],
                  lib_loc=7.0, lib_scale=0.4, lam_scale=2.0, bias_scale=3.0, sigma_scale=0.7, scale_factor=0.5,
                  epsilon=1e-5)
    o = O.OracleDRNBVAE(n_input=40, n_classes=3, n_drugs=2, n_conditions=[4, 5], n_latent=6, layers=[16, 8])
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    for attr in ("encoder", "decoder", "lmb_list", "epsilon", "l_loc", "l_scale", "sigma_scale", "n_latent",
                 "n_classes", "scale_factor", "model", "guide"):
        assert hasattr(m, attr), attr
    assert float(m.l_loc) == 7.0 and float(m.l_scale) == pytest.approx(0.4) and float(m.epsilon) == pytest.approx(1e-5)
    # decoder gets the reversed layers (reference nbvae.py:42)
    assert m.decoder.fc[2].out_features == 8 and m.decoder.fc[5].out_features == 16
    assert m.encoder.fc[2].out_features == 16 and m.encoder.fc[8].out_features == 2 * 6 + 2
    n = D.NBVAE(n_input=40)
    assert [n.encoder.fc[i].out_features for i in (2, 5, 8)] == [64, 64, 64] and n.n_latent == 8
    with pytest.raises(TypeError):
        D.NBVAE(40)  # keyword-only, like the reference


@pytest.mark.skipif(not os.path.exists(O.REFERENCE_MODULES_PY), reason="reference not mounted")
def test_state_dict_interchange_with_reference_modules():
    ref = O.load_reference_modules()
    r_enc, r_dec = ref.Encoder(30, 4, [12, 10]), ref.NBDecoder(4, 30, [10, 12])
    m = D.NBVAE(n_input=30, n_latent=4, layers=[12, 10])
    m.encoder.load_state_dict(r_enc.state_dict())
    m.decoder.load_state_dict(r_dec.state_dict())
    assert torch.equal(m.encoder.fc[2].weight, r_enc.fc[2].weight)


def test_no_cpu_fallback():
    m = D.NBVAE(n_input=20, n_latent=3, layers=[8, 8])
    elbo = D.FusedTraceELBO()
    x = torch.poisson(torch.rand(16, 20) * 3)
    with pytest.raises(RuntimeError, match="CUDA"):
        elbo.loss_and_grads(m.model, m.guide, x)
    with pytest.raises(RuntimeError, match="CUDA"):
        m.encoder(x)
    with pytest.raises(NotImplementedError):
        D.Encoder(20, 3, [8])(x)  # stand-alone sub-module: no owner


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from drvish_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from drvish_b200.parallel import allreduce_grads_and_loss, shard_rows

    cfg = dict(n_input=40, n_latent=3, layers=[16, 16], batch=32, drugs=0)
    model = O.build_model(cfg).double()
    x = O.synthetic_counts(32, 40, seed=11).double()
    noise = {k: v.double() for k, v in O.make_noise(model, 32).items()}
    sl = shard_rows(32, rank, world)
    loss, grads, _ = O.step(model, x[sl], {k: v[sl] for k, v in noise.items()})
    names = list(grads)
    n = sum(g.numel() for g in grads.values())
    flat = torch.zeros(n + 32, dtype=torch.float64)
    flat[:n] = torch.cat([grads[k].reshape(-1) for k in names])
    allreduce_grads_and_loss(flat, n, torch.tensor(loss, dtype=torch.float64))
    # the peer-memory all-reduce is NVLink/CUDA-IPC only: under gloo the step keeps the collective above
    from drvish_b200.parallel import PeerAllReduce

    assert PeerAllReduce.create(n + 32, "cpu") is None
    torch.save({"flat": flat, "n": n, "names": names}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_data_parallel_allreduce_is_sum_over_shards(tmp_path):
    """R-rank result == sum_r oracle(shard_r), loss in the tail slot, ONE collective."""
    world, port = 2, _free_port()
    mp.spawn(_dp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(r0["flat"], r1["flat"])  # replicas agree after the all-reduce
    cfg = dict(n_input=40, n_latent=3, layers=[16, 16], batch=32, drugs=0)
    model = O.build_model(cfg).double()
    x = O.synthetic_counts(32, 40, seed=11).double()
    noise = {k: v.double() for k, v in O.make_noise(model, 32).items()}
    tot_loss, tot = 0.0, None
    for r in range(world):
        sl = slice(16 * r, 16 * (r + 1))
        loss, grads, _ = O.step(model, x[sl], {k: v[sl] for k, v in noise.items()})
        g = torch.cat([grads[k].reshape(-1) for k in r0["names"]])
        tot = g if tot is None else tot + g
        tot_loss += loss
    n = r0["n"]
    torch.testing.assert_close(r0["flat"][:n], tot, rtol=1e-12, atol=1e-12)
    assert abs(float(r0["flat"][n]) - tot_loss) <= 1e-9 * abs(tot_loss)


def test_packed_counts_round_trip_and_rejection():
    """Lossless 16-bit host staging (drv_host_pack_counts_u16): exact for integer counts in [0, 65535],
    refused (first offending index found) otherwise."""
    import ctypes as C

    from drvish_b200 import _lib
    from drvish_b200.data import PackedCounts, pack_batches

    g = torch.Generator().manual_seed(3)
    x = torch.poisson(torch.full((333, 257), 3.0), generator=g)
    x[0, 0], x[5, 7] = 65535.0, 0.0
    for nt in (1, 3, 8):
        p = PackedCounts.pack(x, n_threads=nt)
        assert p is not None and p.shape == (333, 257)
        assert torch.equal(p.unpack_host(), x)
    lib = _lib.load()
    out = torch.empty(x.numel(), dtype=torch.int16)
    for bad in (0.5, -1.0, 65536.0, float("nan"), float("inf")):
        y = x.clone()
        y[200, 100] = bad
        assert PackedCounts.pack(y) is None
        idx = lib.drv_host_pack_counts_u16(C.c_void_p(y.data_ptr()), C.c_void_p(out.data_ptr()), y.numel(), 4)
        assert idx == 200 * 257 + 100
    big = torch.zeros(1 << 20)
    big[70000], big[900000] = 0.25, 0.25
    out = torch.empty(big.numel(), dtype=torch.int16)
    assert lib.drv_host_pack_counts_u16(C.c_void_p(big.data_ptr()), C.c_void_p(out.data_ptr()), big.numel(), 8) == 70000
    sets = pack_batches([[x, "labels"], [x + 0.5, "labels"]])
    assert isinstance(sets[0][0], PackedCounts) and torch.is_tensor(sets[1][0]) and sets[0][1] == "labels"


def test_loaders_have_no_cpu_path():
    """The device loaders (molecule split, CSR gather) raise on host tensors instead of falling back."""
    from drvish_b200.data import MCVDataLoader, split_molecules

    umis = torch.ones(4, 8)
    s = torch.full((4, 1), 0.9)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        split_molecules(umis, s, 1 - s, torch.zeros(4, 1))
    # same constructor surface as the reference's DataLoader subclass (train/mcv_train.py:13-14)
    ds = torch.utils.data.TensorDataset(umis, s, 1 - s, torch.zeros(4, 1))
    dl = MCVDataLoader(ds, batch_size=2, shuffle=False)
    assert len(dl) == 2 and callable(MCVDataLoader.split_molecules)


def test_peer_allreduce_needs_an_nccl_group():
    from drvish_b200.parallel import PeerAllReduce

    assert PeerAllReduce.create(1024, "cpu") is None  # torch.distributed not initialised: NCCL / single rank path


def test_host_pack_u8_with_escapes_is_lossless():
    """drv_host_pack_counts_u8 (host side of the default count-tile transport): bytes + escape list reproduce ANY fp32
    tile bit for bit - small integers as bytes, everything else (>= 255, non-integer, negative, inf) as escapes;
    every vector width's tail, several thread counts; too many escapes -> refused."""
    import torch

    from drvish_b200.data import PackedCounts8

    g = torch.Generator().manual_seed(3)
    for n in (1, 15, 63, 64, 65, 257, 4099, 300_001):
        x = torch.poisson(torch.full((n,), 2.5), generator=g)
        x[n // 2] = 255.0
        if n > 20:
            x[3], x[7], x[11], x[n - 1] = 1.5, -1.0, 70000.0, float("inf")
        for nt in (1, 3):
            p = PackedCounts8.pack(x, n_threads=nt)
            assert p is not None and p.n_esc == (5 if n > 20 else 1), (n, nt, p.n_esc)
            assert torch.equal(p.unpack_host(), x), (n, nt)
    dense = torch.full((100_000,), 1000.0)  # every entry an escape: over the cap
    assert PackedCounts8.pack(dense, n_threads=2) is None


def test_pyro_fused_scheduler_drives_train_until_plateau():
    """The reference's train_until_plateau (train/__init__.py:89-108) restated verbatim around a stub SVI: it calls
    scheduler.step(epoch) and looks at scheduler.optim_objs[*].T_cur / T_0 to detect the end of a cosine cycle."""
    import itertools

    import numpy as np

    import drvish_b200 as D

    model = D.NBVAE(n_input=16, n_latent=2, layers=[8, 8])
    sched = D.PyroFusedScheduler(model, {"optimizer": None, "optim_args": {"lr": 5e-4, "betas": [0.0, 0.9, 0.99], "nesterov": True},
                                         "T_0": 4, "T_mult": 1}, {"clip_norm": 20.0})
    assert sched.optim.param_groups[0]["lr"] == pytest.approx(5e-4 / 3)  # aggmo.py:22

    class StubSVI:  # losses improve for two cycles, then stall
        def __init__(self):
            self.k = 0

        def step(self, *xs):
            return 100.0

        def evaluate_loss(self, *xs):
            self.k += 1
            return max(50.0, 100.0 - 5.0 * self.k)

    svi, train_loss, val_loss = StubSVI(), [], []
    best, threshold, min_cycles, cycle = np.inf, 0.01, 3, 0
    lrs = []
    for epoch in itertools.count():
        train_loss.append(svi.step())
        val_loss.append(svi.evaluate_loss())
        sched.step(epoch)
        lrs.append(sched.optim.param_groups[0]["lr"])
        if any(opt.T_cur == opt.T_0 - 1 for opt in sched.optim_objs.values()):
            cycle += 1
            if 0 <= val_loss[-1] < best * (1.0 - threshold):
                best = val_loss[-1]
            elif cycle >= min_cycles:
                break
        assert epoch < 200
    assert cycle >= min_cycles and len(train_loss) == len(val_loss) == epoch + 1
    assert max(lrs) == pytest.approx(5e-4 / 3, rel=0.2) and min(lrs) < 0.3 * max(lrs)  # the cosine really moves the fused optimizer's lr
    state = sched.get_state()
    sched.set_state(state)


def test_hybrid_transport_split_balances_bus_and_packer():
    """PinnedBatchPrefetcher ships a fp32 host tile over two paths at once (data.py:_stage_hybrid); the split must
    make the bus time - fp32 rows AND packed rows - equal to the packing time, and pack everything when the bus is
    slower than the packer's output."""
    from drvish_b200.data import hybrid_fp32_fraction as hf

    for a, b, c in ((45e9, 110e9, 0.25), (45e9, 70e9, 0.26), (27e9, 110e9, 0.25), (55e9, 20e9, 0.5), (10e9, 10e9, 1.0)):
        f = hf(a, b, c)
        assert 0.0 <= f <= 1.0
        bus, pack = (f + (1 - f) * c) / a, (1 - f) / b
        if f > 0.0:
            assert abs(bus - pack) <= 1e-9 * max(bus, pack), (a, b, c, f)
        else:
            assert c * b >= a  # the packed stream alone saturates the bus
    assert hf(45e9, 110e9, 0.25) == pytest.approx(17.5 / 127.5)
    assert hf(0.0, 1.0, 0.25) == 0.0 and hf(1.0, 0.0, 0.25) == 1.0
