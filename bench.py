#!/usr/bin/env python
"""Headline benchmark: cells/sec per SVI step (ELBO forward + backward) - BASELINE.json `metric`.

    python bench.py --gpus N --steps K --warmup W              # this framework (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...    # reference CPU path on the host cores

Workload (SURVEY 8d / BASELINE.md 4): config 3 by default - synthetic NB counts, 100k cells x
5k genes, latent 10, layers [256,256], 8 drugs x 8 doses, K=8 classes, minibatch 4096 cells per
GPU (weak scaling: every rank steps its own 4096-cell shard, gradients summed with one NCCL
all-reduce per step).  One "step" = noise -> ELBO -> every parameter gradient (-> all-reduce);
the optimizer update is excluded (it is not part of the north star).  Inputs rotate over >= 20
distinct resident minibatches (1.6 GB > the 126 MB L2), so no L2 flush is needed.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = same metric through the public API (FusedTraceELBO.loss_and_grads, what pyro's SVI.step
calls) with pinned HOST minibatches copied H2D and the loss read back D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    2: dict(name="cfg2: 10k cells x 2k genes, latent 10, layers [128,128], NBVAE", n_cells=10_000,
            n_input=2000, n_latent=10, layers=[128, 128], batch=4096, drugs=0, doses=0, classes=0),
    3: dict(name="cfg3: 100k cells x 5k genes, latent 10, layers [256,256], 8 drugs x 8 doses, K=8, DRNBVAE",
            n_cells=100_000, n_input=5000, n_latent=10, layers=[256, 256], batch=4096, drugs=8, doses=8, classes=8),
    4: dict(name="cfg4: 1M cells x 20k genes, latent 32, layers [512,512], NBVAE", n_cells=1_000_000,
            n_input=20000, n_latent=32, layers=[512, 512], batch=4096, drugs=0, doses=0, classes=0),
    5: dict(name="cfg5: 4M cells x 30k genes, latent 64, layers [512,512], 96 drugs x 10 doses, K=8, DRNBVAE",
            n_cells=4_000_000, n_input=30000, n_latent=64, layers=[512, 512], batch=4096, drugs=96, doses=10, classes=8),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tensor=p.get("bf16_tflops_sustained", p["bf16_tflops"]), which="measured")
    return dict(hbm=6650.0, tensor=1400.0, which="fallback")


def decoder_algorithmic(cfg):
    """bytes_alg(dec) = 8BG + 8Bh + 16Gh ; flops_alg(dec) = 12 B G h."""
    B, G, h = cfg["batch"], cfg["n_input"], cfg["layers"][0]
    return 8.0 * B * G + 8.0 * B * h + 16.0 * G * h, 12.0 * B * G * h


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def finish(self):
        self._halt.set()
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def synth_batches(cfg, n_batches, seed, device):
    """x ~ NB(2, logits=u_b+v_g), generated on the device; labels = b mod K; y ~ N(0,1)."""
    g = torch.Generator(device=device).manual_seed(seed)
    B, G = cfg["batch"], cfg["n_input"]
    v = -1.0 + torch.randn(1, G, generator=g, device=device)
    xs = []
    for _ in range(n_batches):
        u = 0.5 * torch.randn(B, 1, generator=g, device=device)
        rate = torch._standard_gamma(torch.full((B, G), 2.0, device=device), generator=g) * torch.exp(u + v)
        xs.append(torch.poisson(rate, generator=g).float())
    labels = y = None
    if cfg["drugs"]:
        labels = (torch.arange(B, device=device) % cfg["classes"]).long()
        y = [torch.randn(cfg["classes"], cfg["doses"], 1, generator=g, device=device) for _ in range(cfg["drugs"])]
    return xs, labels, y


def build_native_model(cfg, device):
    import drvish_b200 as D

    torch.manual_seed(0)
    kw = dict(n_input=cfg["n_input"], n_latent=cfg["n_latent"], layers=cfg["layers"])
    if cfg["drugs"]:
        m = D.DRNBVAE(n_classes=cfg["classes"], n_drugs=cfg["drugs"], n_conditions=[cfg["doses"]] * cfg["drugs"], **kw)
    else:
        m = D.NBVAE(**kw)
    return m.to(device)


def oracle_cpu_rate(cfg, steps, warmup, batch=None):
    """The reference's CPU implementation of the path: the oracle's PyTorch-eager restatement
    of the Pyro step (reference Encoder/NBDecoder/logit_mean_sigmoid arithmetic +
    torch.distributions + .backward()), fp32, all host threads.  Pyro itself cannot be installed
    (no wheel, no network), so its ~1-2 ms/step of tracing overhead is absent."""
    from oracle import drvish_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ocfg = dict(cfg)
    B = batch or cfg["batch"]
    ocfg["batch"] = B
    model = O.build_model(ocfg, seed=0)
    x = O.synthetic_counts(B, cfg["n_input"], seed=1234)
    noise = O.make_noise(model, B, seed=1)
    labels = y = None
    if cfg["drugs"]:
        labels = torch.arange(B) % cfg["classes"]
        y = [torch.randn(cfg["classes"], cfg["doses"], 1) for _ in range(cfg["drugs"])]
    for _ in range(warmup):
        O.step(model, x, noise, labels, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.step(model, x, noise, labels, y)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3, cores, B


def config_dict(cfg, world, n_batches, cuda_graph=True):
    """`config` of the JSON line - the SAME keys and values for both arms (the driver compares them)."""
    B, G = cfg["batch"], cfg["n_input"]
    return {"workload": cfg["name"], "minibatch_per_gpu": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "l2": f"inputs rotate over {n_batches} resident minibatches ({n_batches * B * G * 4 / 1e9:.2f} GB > 126 MB L2), no flush",
            "optimizer": "excluded from value/e2e (see with_optimizer)", "noise": "Philox in-kernel draws each step",
            "cuda_graph": cuda_graph,
            "all_reduce": "none (1 GPU)" if world == 1 else "one all-reduce of the flat gradient buffer per step (see `collective`)"}


def run_reference(args, cfg):
    """Reference arm: the reference's CPU implementation of the path (the oracle port - the reference itself cannot be
    installed: pyro-ppl has no wheel here) on all host cores, same config / steps / warmup as the native arm; every step
    is one 4096-cell minibatch of the workload.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, ms, cores, B = oracle_cpu_rate(cfg, args.steps, args.warmup)
    sample = f"{args.steps} timed steps ({args.warmup} warm-up) of one {B}-cell minibatch of {cfg['name']}"
    line = {
        "impl": "reference", "metric": "cells/sec per SVI step (ELBO fwd+bwd)", "value": rate, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfg, args.gpus, max(20, args.batches), not args.no_graph),
        "cpu_baseline": {"value": rate, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "device": "host CPU (PyTorch-eager restatement of the Pyro step, oracle/drvish_oracle.py)",
    }
    print(json.dumps(line), flush=True)


MUFU_PER_ELEMENT = {"k_dec<1>": 1, "k_dec<2>": 8, "k_dec<3>": 1}  # MUFU (ex2 / lg2 / rcp) issued per (cell, gene) element


def source_hash():
    """Hash of the kernel sources the committed ncu traffic figures belong to (profiles/ncu_traffic.json is only
    used while it matches: a stale capture must not be reported as this build's traffic)."""
    import hashlib

    h = hashlib.sha256()
    for f in ("tc_decoder.cu", "tc_gemm.cu", "tc_common.cuh", "common.cuh"):
        with open(os.path.join(ROOT, "drvish_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def bound_times(bytes_, flops, mufu, peaks, sm_mhz):
    """Lower bounds (seconds) of a piece of work from the three rooflines that can bind it: HBM bytes, tensor-core
    flops, SFU (MUFU) operations at 16 per clock and SM."""
    return {"hbm": bytes_ / (peaks["hbm"] * 1e9), "tensor": flops / (peaks["tensor"] * 1e12),
            "sfu": mufu / (148 * 16 * sm_mhz * 1e6)}


def roof_entry(ms, bytes_, flops, mufu, peaks, sm_mhz):
    t = bound_times(bytes_, flops, mufu, peaks, sm_mhz)
    bound = max(t, key=t.get)
    work, peak, unit = {"hbm": (bytes_ / 1e9, peaks["hbm"], "GB/s"), "tensor": (flops / 1e12, peaks["tensor"], "TFLOP/s"),
                        "sfu": (mufu / 1e9, 148 * 16 * sm_mhz * 1e-3, "GMUFU/s")}[bound]
    ach = work / (ms * 1e-3)
    return {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
            "bound_us": {k: round(v * 1e6, 2) for k, v in t.items()}}


class Workload:
    """One bench config on this rank: model, resident synthetic minibatches, the fused ELBO."""

    def __init__(self, args, cfg, device, world, rank):
        import drvish_b200 as D

        self.args, self.cfg, self.device, self.world, self.rank = args, cfg, device, world, rank
        self.model = build_native_model(cfg, device)
        self.rt = self.model._runtime()
        self.elbo = D.FusedTraceELBO(seed=1234, data_parallel=(world > 1), use_cuda_graph=not args.no_graph)
        self.n_batches = max(20, args.batches)
        self.xs, self.labels, self.y = synth_batches(cfg, self.n_batches, 1234 + 3 + rank, device)
        self.yflat = self.elbo._flatten_y(self.rt, self.y, device) if cfg["drugs"] else None
        self.rt.ensure_flat()
        if world > 1 and not self.elbo.dp_overlap:
            self.elbo.dp_prepare(self.rt)  # peer-mapped gradient buffer (before any graph is captured)

    def barrier(self):
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        import torch.distributed as dist

        t = torch.tensor([v], device=self.device, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_step(self, i, single_call=False, reduce=True):
        """exactly what FusedTraceELBO.loss_and_grads enqueues, minus the per-step loss read-back"""
        import torch.distributed as dist

        e, rt = self.elbo, self.rt
        xb = self.xs[i % self.n_batches]
        if self.world > 1 and not single_call and not e.dp_overlap:
            e.dp_step(rt, xb, self.labels, self.yflat, reduce=reduce)
        elif self.world > 1 and not single_call:
            e._enqueue(rt, xb, self.labels, self.yflat, True, part=1)
            w1 = dist.all_reduce(rt.flat_grads[rt.decoder_offset:], op=dist.ReduceOp.SUM, async_op=True)
            e._enqueue(rt, xb, self.labels, self.yflat, True, part=2)
            w2 = dist.all_reduce(rt.flat_grads[:rt.decoder_offset], op=dist.ReduceOp.SUM, async_op=True)
            w1.wait()
            w2.wait()
        else:
            e._enqueue(rt, xb, self.labels, self.yflat, True)
        e.step_index += 1

    def per_rank(self, v):
        import torch.distributed as dist

        if self.world == 1:
            return [v]
        t = torch.tensor([v], device=self.device, dtype=torch.float64)
        out = torch.empty(self.world, device=self.device, dtype=torch.float64)
        dist.all_gather_into_tensor(out, t)
        return [round(float(x), 5) for x in out.tolist()]

    def timed_local(self, steps, warmup):
        """The step WITHOUT its collective (same CUDA-graph treatment as the full step), per rank: what every rank's
        own work takes when nobody waits for anybody - the rest of ms_per_step is the collective plus the wait for
        the slowest rank."""
        fn = lambda i: self.device_step(i, reduce=False)
        if not self.args.no_graph:
            for i in range(2 * self.n_batches):
                fn(i)
        for i in range(warmup):
            fn(i)
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            fn(warmup + i)
        ev1.record()
        self.barrier()
        mine = ev0.elapsed_time(ev1) / steps
        return self.max_over_ranks(mine), self.per_rank(mine)

    def step_jitter(self, steps=40):
        """Per-step durations of the step WITHOUT its collective on every rank (one event pair per step, a device sync
        between steps): mean, standard deviation and the mean over steps of (slowest rank - mean rank) - the wait a
        blocking collective adds on top of its own duration."""
        import torch.distributed as dist

        fn = lambda i: self.device_step(i, reduce=False)
        for i in range(3):
            fn(i)
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for i, (a, b) in enumerate(evs):
            a.record()
            fn(3 + i)
            b.record()
        torch.cuda.synchronize()
        mine = torch.tensor([a.elapsed_time(b) for a, b in evs], device=self.device, dtype=torch.float64)
        allr = torch.empty(self.world * steps, device=self.device, dtype=torch.float64)
        if self.world > 1:
            dist.all_gather_into_tensor(allr, mine)
        else:
            allr.copy_(mine)
        m = allr.view(self.world, steps)
        return {"mean_ms": float(m.mean()), "std_ms_per_rank": [round(float(v), 4) for v in m.std(dim=1).tolist()],
                "mean_slowest_minus_mean_ms": float((m.max(dim=0).values - m.mean(dim=0)).mean()),
                "note": "back-to-back graph replays, one event pair per step"}

    def timed(self, steps, warmup, fn=None, sample_clocks=False):
        """W untimed + K timed steps between barriers, CUDA events on the launching stream, max over ranks (ms/step)."""
        fn = fn or self.device_step
        for i in range(warmup):
            fn(i)
        self.barrier()
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(self.device.index)
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            fn(warmup + i)
        ev1.record()
        self.barrier()
        clocks = sampler.finish() if sampler else None
        return self.max_over_ranks(ev0.elapsed_time(ev1)) / steps, clocks

    def capture_graphs(self):
        if not self.args.no_graph:
            for i in range(2 * self.n_batches):  # untimed: every resident minibatch is seen twice -> its graph is captured
                self.device_step(i)

    def dp_check(self):
        """Data-parallel gradients checked against an independent reduction.  (1) The collective: every rank's
        PRE-reduction flat gradient buffer is all-gathered through NCCL and summed in fp64; the step's own collective,
        applied to the same buffer, must reproduce it (<= 1e-6) with bit-identical results on all ranks.  (2) The step as
        it is timed (collective launched by the step itself, decoder slice overlapped with the encoder backward): the
        same minibatch and the same Philox counter once without and once with the collective; the reduced buffer
        must equal the gathered sum of the un-reduced run up to the run-to-run noise of the atomics (<= 1e-5)."""
        import torch.distributed as dist

        rt, e, n = self.rt, self.elbo, self.rt.n_param_floats + 1

        def gathered_sum():
            local = rt.flat_grads[:n].clone()
            gathered = torch.empty(self.world * n, dtype=torch.float32, device=self.device)
            dist.all_gather_into_tensor(gathered, local)
            return gathered.view(self.world, n).double().sum(0)

        def errs(ref):
            got = rt.flat_grads[:n].double()
            bits = rt.flat_grads[:n].view(torch.int32).to(torch.int64)
            chk = torch.stack([bits.sum(), (bits * torch.arange(1, n + 1, device=self.device) % 1000003).sum()])
            allchk = torch.empty(self.world * 2, dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(allchk, chk)
            same = bool((allchk.view(self.world, 2) == allchk.view(self.world, 2)[0]).all().item())
            return (float(((got - ref).norm() / ref.norm()).item()), float(((got - ref).abs().max() / ref.abs().max()).item()), same)

        ctr = e._step_ctr.clone() if e._step_ctr is not None else None
        self.device_step(0, reduce=False)
        torch.cuda.synchronize()
        ref = gathered_sum()
        e.dp_all_reduce(rt)
        torch.cuda.synchronize()
        n1, m1, same1 = errs(ref)
        out = {"max_rel_err": max(n1, m1), "norm_rel_err": n1, "bit_identical_across_ranks": same1,
               "collective": self.collective(), "floats": n, "tolerance": 1e-6}
        if ctr is not None:
            e._step_ctr.copy_(ctr)  # same noise again
            self.device_step(0, reduce=True)
            torch.cuda.synchronize()
            n2, m2, same2 = errs(ref)
            out["timed_step"] = {"norm_rel_err": n2, "max_rel_err": max(n2, m2), "bit_identical_across_ranks": same2,
                                 "tolerance": 1e-5, "fused_in_step": bool(e.dp_fused)}
        return out

    def collective(self):
        if self.world == 1:
            return "none (1 GPU)"
        if self.elbo.dp_overlap:
            return "two-part step, NCCL all-reduces overlapping the encoder backward"
        if self.elbo._peer.get(id(self.rt)) is not None:
            if self.elbo.dp_fused and not self.rt.shape_padded(self.cfg["batch"]):
                return ("k_peer_allreduce over NVLink peer memory, launched by the step itself inside its CUDA graph: slice "
                        "[last decoder weight .. loss/status tail] on a comm stream (32 CTAs) next to the trunk backward, the "
                        "rest at the end of the step (k_comm.cu, DRV_FLAG_FUSED_ALLREDUCE)")
            return "drv_peer_allreduce: one kernel over NVLink peer memory per step (k_comm.cu)"
        return "one NCCL all-reduce per step"

    def allreduce_alone_ms(self, iters=20):
        """The step's collective timed back to back on the flat gradient buffer (not overlapped with anything)."""
        if self.world == 1:
            return 0.0
        ms, _ = self.timed(iters, 3, fn=lambda i: self.elbo.dp_all_reduce(self.rt))
        return ms

    def profile_sections(self, lib, n_prof):
        """Live kernel timing for the roofline: CUDA events recorded by the library on the step's own stream, around
        the sections and around the single kernels of the fused decoder path.  The side stream is switched off
        (flag 64) so that a section contains ALL of its work; direct launches (no graph)."""
        os.environ["DRVISH_B200_DEBUG_FLAGS"] = "64"
        graph = self.elbo.use_cuda_graph
        self.elbo.use_cuda_graph = False
        for i in range(3):
            self.device_step(i, single_call=True)
        torch.cuda.synchronize()
        lib.drv_profile_enable(1)
        for i in range(n_prof):
            self.device_step(i, single_call=True)
        torch.cuda.synchronize()
        sec, kms, nprof = (C.c_float * 7)(), (C.c_float * 5)(), C.c_int()
        lib.drv_profile_read(sec, C.byref(nprof))
        lib.drv_profile_read_kernels(kms)
        lib.drv_profile_enable(0)
        os.environ["DRVISH_B200_DEBUG_FLAGS"] = "0"
        self.elbo.use_cuda_graph = graph
        return [float(v) for v in sec], [float(v) for v in kms], int(nprof.value)

    def roofline(self, sec, kms, nprof, peaks, sm_mhz, config_id):
        """Roofline of the fused decoder-last-layer + NB-likelihood step: algorithmic work / measured time
        against the slowest of the HBM, tensor-core and SFU bounds."""
        cfg = self.cfg
        Bf, Gf, hf = float(cfg["batch"]), float(cfg["n_input"]), float(cfg["layers"][0])
        bytes_alg, flops_alg = decoder_algorithmic(cfg)
        mufu = sum(MUFU_PER_ELEMENT.values()) * Bf * Gf
        t_dec_ms = sec[2] + sec[4]
        roof = roof_entry(t_dec_ms, bytes_alg, flops_alg, mufu, peaks, sm_mhz)
        traffic, tnote = {}, "no ncu capture committed for this config"
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("source_hash") == source_hash():
                traffic = tj.get(f"cfg{config_id}", {})
                tnote = tj.get("source", "")
            else:
                tnote = "profiles/ncu_traffic.json belongs to other kernel sources (hash mismatch): not reported"
        roof["traffic"] = traffic.get("decoder_step_bytes")
        roof["traffic_source"] = tnote
        roof["kernel"] = ("fused decoder step = k_act + k_dec<1> (logsumexp) + k_dec<2> (NB likelihood + NB part of dlogits) + "
                          "k_dec3w (softmax coupling + the scale half's weight gradient, d logits consumed in shared memory; "
                          "h > 256: k_dec<3> + dW GEMM) + dAct GEMM + dW GEMM of the log_r half (sections DEC_NB_FWD + DEC_NB_BWD, "
                          "side stream off)")
        roof["ms"] = round(t_dec_ms, 4)
        roof["algorithmic"] = {"bytes": bytes_alg, "flops": flops_alg, "mufu_ops": mufu,
                               "formula": "8BG+8Bh+16Gh bytes, 12BGh flops (SURVEY 8d); MUFU = (1 + 8 + 1) B G as issued by "
                                          "phases 1/2/3 (2 ex2 + 4 lg2 + 2 rcp per element in the NB block, 1 ex2 each in the "
                                          "softmax passes), 16 MUFU per clock and SM at the SM clock seen under load"}
        roof["peak_source"] = peaks["which"] + (" (HBM: copy bandwidth; tensor: bf16 sustained cuBLAS - the decoder kernels use "
                                                "kind::f16, fp16 operands with fp32 accumulation, whose nominal peak is the bf16 one; "
                                                "SFU: 148 SMs x 16 lanes x SM clock)")
        roof["section_ms"] = {k: round(sec[i], 4) for i, k in enumerate(
            ["enc_fwd", "dec_trunk_fwd", "dec_nb_fwd", "dose_response", "dec_nb_bwd", "dec_trunk_bwd", "enc_bwd"])}
        roof["profiled_steps"] = nprof
        # per-kernel entries: the bytes / flops / MUFU each launch must move TODAY (fp16 operands and the fp16 dlogits
        # hand-off between the kernels included - that hand-off is not algorithmic, which is why only the whole-step
        # entry above is the roofline claim) over its own event-timed duration
        # h <= 256: phase 3 and the weight gradient of the scale half are ONE kernel (k_dec3w: the finished ds tile is an
        # operand of the dW MMA while it is still in shared memory); the separate dW GEMM only covers the log_r half
        fused_dw = hf <= 256 and os.environ.get("DRV_DEC_FUSED_DW", "1") != "0"
        kdefs = [
            ("k_dec<1> logsumexp", 2 * Gf * hf + 2 * Bf * hf, 2 * Bf * Gf * hf, 1 * Bf * Gf),
            ("k_dec<2> NB likelihood (+ NB part of dlogits)", 4 * Bf * Gf + 4 * Bf * Gf + 4 * Gf * hf + 2 * Bf * hf, 4 * Bf * Gf * hf, 8 * Bf * Gf),
        ]
        if fused_dw:
            kdefs += [
                ("k_dec3w softmax coupling + dW (scale half) in one kernel", 2 * Bf * Gf + 2 * Bf * Gf + 2 * Gf * hf + 2 * Bf * hf + 4 * Gf * hf,
                 4 * Bf * Gf * hf, 1 * Bf * Gf),
                ("dAct GEMM", 4 * Bf * Gf + 4 * Gf * hf + 4 * Bf * hf, 4 * Bf * Gf * hf, 0.0),
                ("dW GEMM (log_r half)", 2 * Bf * Gf + 2 * Bf * hf + 4 * Gf * hf, 2 * Bf * Gf * hf, 0.0),
            ]
        else:
            kdefs += [
                ("k_dec<3> softmax coupling", 2 * Bf * Gf + 2 * Bf * Gf + 2 * Gf * hf + 2 * Bf * hf, 2 * Bf * Gf * hf, 1 * Bf * Gf),
                ("dAct GEMM", 4 * Bf * Gf + 4 * Gf * hf + 4 * Bf * hf, 4 * Bf * Gf * hf, 0.0),
                ("dW GEMM", 4 * Bf * Gf + 2 * Bf * hf + 8 * Gf * hf, 4 * Bf * Gf * hf, 0.0),
            ]
        rk = []
        for i, (name, kb, kf, km) in enumerate(kdefs):
            if kms[i] <= 0:
                continue
            ent = {"kernel": name, "ms": round(kms[i], 4)}
            ent.update(roof_entry(kms[i], kb, kf, km, peaks, sm_mhz))
            ent["traffic"] = traffic.get(name.split(" ")[0])
            rk.append(ent)
        roof["kernels"] = rk
        return roof


def cpu_baseline(cfg, steps, batch=None):
    rate, ms, cores, Bc = oracle_cpu_rate(cfg, steps, 2, batch=batch)
    return {"value": rate, "unit": "cells/s", "cores": cores, "kind": "port",
            "sample": f"{steps} timed steps (2 warm-up) of one {Bc}-cell minibatch, fp32, PyTorch-eager restatement of the "
                      "Pyro step (oracle/drvish_oracle.py)", "ms_per_step": ms}


def measure_other_config(args, config_id, device, world, rank, lib, peaks):
    """A short device-timed run of another BASELINE.json config in the same process (value, ms/step, all-reduce,
    decoder roofline, CPU baseline on a 512-cell sample)."""
    cfg = CONFIGS[config_id]
    w = Workload(args, cfg, device, world, rank)
    w.capture_graphs()
    steps = max(10, min(args.steps, 30))
    ms, clocks = w.timed(steps, 3, sample_clocks=True)
    ent = {"workload": cfg["name"], "n_gpus": world, "steps": steps, "ms_per_step": ms,
           "value": world * cfg["batch"] / (ms * 1e-3), "unit": "cells/s", "launches_per_step": w.rt.last_launches + 1,
           "grad_bytes": 4 * w.rt.n_param_floats, "clocks": clocks}
    if world > 1:
        ms_local, per_rank = w.timed_local(steps, 3)
        ent["all_reduce"] = {"collective": w.collective(), "alone_ms": w.allreduce_alone_ms(),
                             "exposed_ms": ms - ms_local, "ms_per_step_without": ms_local, "per_rank_ms_without": per_rank}
        ent["dp_check"] = w.dp_check()
    sec, kms, nprof = w.profile_sections(lib, min(steps, 16))
    sm_mhz = (clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    if rank == 0:
        roof = w.roofline(sec, kms, nprof, peaks, sm_mhz, config_id)
        ent["roofline"] = {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "bound_us", "ms", "traffic")}
        ent["roofline"]["kernels_ms"] = {k["kernel"].split(" ")[0]: k["ms"] for k in roof["kernels"]}
        if not args.no_cpu_baseline:
            ent["cpu_baseline"] = cpu_baseline(cfg, 3, batch=512)
    del w
    torch.cuda.empty_cache()
    return ent


def run_native(args, cfg):
    import torch.distributed as dist

    from drvish_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the native path has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()
    peaks = load_peaks()
    w = Workload(args, cfg, device, world, rank)
    model, rt, elbo, xs, labels, y = w.model, w.rt, w.elbo, w.xs, w.labels, w.y
    n_batches = w.n_batches
    B, G = cfg["batch"], cfg["n_input"]
    barrier = w.barrier

    # ---- device-resident timing (`value`) ----------------------------------------------------
    # same code path as loss_and_grads minus the per-step loss read-back (one sync at the end)
    w.capture_graphs()
    ms_per_step, clocks = w.timed(args.steps, args.warmup, sample_clocks=True)
    launches_per_step = rt.last_launches + 1  # + the noise draw (one launch)
    # (the data-parallel step counts its all-reduce kernels itself when it launches them: DRV_FLAG_FUSED_ALLREDUCE)
    loss_last, status = rt.read_out()
    value = world * B / (ms_per_step * 1e-3)

    # ---- data parallel: correctness of the collective + its cost ------------------------------
    dp_check = all_reduce = None
    if world > 1 and not elbo.dp_overlap:
        dp_check = w.dp_check()
        ms_local, per_rank = w.timed_local(args.steps, 3)
        all_reduce = {"collective": w.collective(), "alone_ms": w.allreduce_alone_ms(), "exposed_ms": ms_per_step - ms_local,
                      "ms_per_step_without": ms_local, "per_rank_ms_without": per_rank, "bytes": 4 * (rt.n_param_floats + 1),
                      "step_jitter": w.step_jitter()}
        ts = dp_check.get("timed_step")
        if dp_check["max_rel_err"] > dp_check["tolerance"] or (ts and ts["norm_rel_err"] > ts["tolerance"]):
            raise SystemExit(f"bench.py: data-parallel gradient check failed: {dp_check}")

    # ---- the same loop with the fused optimizer step behind every ELBO step (AggMo, 3 betas,
    # nesterov, per-tensor clip_norm 20 as in the reference notebook); an extra, the headline excludes it
    import drvish_b200 as D

    opt = D.FusedAggMo(model, lr=5e-4, betas=[0.0, 0.9, 0.99], nesterov=True, clip_norm=20.0)
    params_before = rt.flat_params.clone()

    def step_opt(i):
        w.device_step(i)
        opt.step()

    ms_opt, _ = w.timed(args.steps, 3, fn=step_opt)
    with_opt = {"value": world * B / (ms_opt * 1e-3), "unit": "cells/s", "ms_per_step": ms_opt,
                "optimizer": "FusedAggMo(betas=[0,0.9,0.99], nesterov, clip_norm=20): drv_aggmo_step, 3-4 launches",
                "optimizer_ms": ms_opt - ms_per_step}
    rt.flat_params.copy_(params_before)  # the measurements below run on the initial weights again

    sec, kms, nprof = w.profile_sections(lib, min(args.steps, 32))

    # ---- end-to-end through the public API, host minibatches --------------------------------
    # every step's inputs start in pinned HOST memory; PinnedBatchPrefetcher issues the H2D copy of
    # batch i+1 on a side stream while step i runs; loss_and_grads reads the loss back (D2H) each step
    from drvish_b200.data import PinnedBatchPrefetcher, pack_batches

    host_x = [xs[i].cpu().pin_memory() for i in range(min(4, n_batches))]
    e2e_steps = max(10, args.steps // 2)
    tail = [labels, y] if cfg["drugs"] else []
    fp32_set = [[hx] + tail for hx in host_x]
    # the same minibatches packed once to 16 bits (dataset-build-time work, lossless: data.PackedCounts)
    u16_set = pack_batches(fp32_set)
    u8_set = pack_batches(fp32_set, kind="u8")  # ... or to one byte per entry + escapes (data.PackedCounts8)

    def run_e2e(pf, dataset, n):
        pf.batches = (dataset[i % len(dataset)] for i in range(n))
        last = None
        for batch in pf:
            last = elbo.loss_and_grads(model.model, model.guide, *batch)
        return last

    def time_e2e(dataset, on_the_fly=None):
        pf = PinnedBatchPrefetcher(None, device, pack_on_the_fly=on_the_fly)
        run_e2e(pf, dataset, 3 * pf.SLOTS)  # untimed: every staging slot gets its CUDA graph
        barrier()
        pf.h2d_bytes, pf.pack_seconds, pf.packed_batches = 0, 0.0, 0
        t0 = time.perf_counter()
        run_e2e(pf, dataset, e2e_steps)
        torch.cuda.synchronize()
        dt = w.max_over_ranks(time.perf_counter() - t0)
        pack_ms = 1e3 * pf.pack_seconds / pf.packed_batches if pf.packed_batches else 0.0
        return world * B * e2e_steps / dt, pf.h2d_bytes // e2e_steps, pack_ms

    # ---- the reference's GPU-resident sparse loader (data/cupy.py): CSR dataset on the device, every
    # batch densified by drv_csr_gather_rows, no host copy of counts (an extra: NOT an e2e number in the
    # contract's sense, its inputs never cross the bus)
    from drvish_b200.data import CsrCountMatrix

    n_csr = min(4, n_batches)
    csr = CsrCountMatrix.from_torch(torch.cat([xs[i] for i in range(n_csr)]).to_sparse_csr(), device=device)
    csr_idx = [torch.arange(i * B, (i + 1) * B, device=device) for i in range(n_csr)]

    def run_csr(n):
        last = None
        for i in range(n):
            last = elbo.loss_and_grads(model.model, model.guide, csr.dense_rows(csr_idx[i % n_csr]), *tail)
        return last

    run_csr(9)
    barrier()
    t0 = time.perf_counter()
    run_csr(e2e_steps)
    torch.cuda.synchronize()
    csr_value = world * B * e2e_steps / w.max_over_ranks(time.perf_counter() - t0)
    csr_nnz = int(csr.data.numel() // n_csr)

    # default transport: fp32 host minibatches packed to one byte per entry (+ escapes) by the prefetcher's worker
    e2e_value, e2e_h2d, e2e_pack_ms = time_e2e(fp32_set, on_the_fly="u8")
    e2e_f32, e2e_f32_h2d, _ = time_e2e(fp32_set, on_the_fly=None)
    e2e_u16, e2e_u16_h2d, _ = time_e2e(u16_set, on_the_fly=None)
    e2e_u8d, e2e_u8d_h2d, _ = time_e2e(u8_set, on_the_fly=None)
    e2e_fly, e2e_fly_h2d, e2e_fly_ms = time_e2e(fp32_set, on_the_fly="u16")
    from drvish_b200.data import default_pack_threads
    pack_isa = lib.drv_host_pack_isa().decode()

    # ---- the other BASELINE.json configs, short runs in the same process: cfg 2 on one GPU, cfgs 4 and 5 where
    # BASELINE.json places them (8 GPUs); --other-configs forces a list
    if args.other_configs is not None:
        others = [int(v) for v in args.other_configs.split(",") if v]
    else:
        others = [2] if world == 1 else ([4, 5] if world == 8 else [])
    others = [c for c in others if c != args.config]
    del host_x, fp32_set, u16_set, u8_set, csr, csr_idx
    other_configs = {}
    if others:
        keep_sm = (clocks or {}).get("sm_mhz")
        del w, model, rt, elbo, xs, labels, y, opt, params_before
        torch.cuda.empty_cache()
        for c in others:
            other_configs[f"cfg{c}"] = measure_other_config(args, c, device, world, rank, lib, peaks)
        w = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    if w is None:  # the main workload was released before the other configs ran: rebuild the roofline from its numbers
        w = Workload.__new__(Workload)
        w.cfg = cfg
    roof = Workload.roofline(w, sec, kms, nprof, peaks, sm_mhz, args.config)

    # ---- CPU baseline on the host cores (bounded sample) ---------------------------------------
    cpu = None if args.no_cpu_baseline else cpu_baseline(cfg, args.cpu_steps)

    line = {
        "metric": "cells/sec per SVI step (ELBO fwd+bwd)", "value": value, "unit": "cells/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 / tf32 tensor-core operands with fp32 accumulation where stated in DESIGN.md)",
        "data": "synthetic",
        "config": config_dict(cfg, world, n_batches, not args.no_graph),
        "collective": all_reduce["collective"] if all_reduce else "none (1 GPU)",
        "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": 16,
                "steps": e2e_steps, "api": "FusedTraceELBO.loss_and_grads(model.model, model.guide, *batch) fed by PinnedBatchPrefetcher: "
                       "every step's fp32 minibatch starts in pinned HOST memory; the prefetcher's worker packs it losslessly to one byte "
                       "per entry + an escape list (drv_host_pack_counts_u8), copies the bytes H2D on a side stream while earlier "
                       "steps run, and the device widens them back to the fp32 tile (drv_unpack_counts_u8); the loss is read back every step",
                "transport": "u8 + escapes, packed on the fly from fp32 host tensors",
                "pack_ms_per_batch": e2e_pack_ms, "pack_threads": default_pack_threads(), "pack_isa": pack_isa,
                "h2d_GBps": e2e_h2d * e2e_value / (world * B) / 1e9,
                # same API, the reference's transport: the fp32 tile itself crosses the bus (PCIe-bound)
                "fp32_transport": {"value": e2e_f32, "unit": "cells/s", "h2d_bytes_per_step": e2e_f32_h2d,
                                   "bound": "PCIe: %.1f GB/s host->device" % (e2e_f32_h2d * e2e_f32 / (world * B) / 1e9)},
                # same API, minibatches packed ONCE when the dataset is built (pack_batches): lossless uint16, or bytes + escapes
                "packed_u16_dataset": {"value": e2e_u16, "unit": "cells/s", "h2d_bytes_per_step": e2e_u16_h2d},
                "packed_u8_dataset": {"value": e2e_u8d, "unit": "cells/s", "h2d_bytes_per_step": e2e_u8d_h2d},
                # same API, fp32 host minibatches packed to uint16 by the prefetcher's worker every step
                "packed_u16_on_the_fly": {"value": e2e_fly, "unit": "cells/s", "h2d_bytes_per_step": e2e_fly_h2d,
                                          "pack_ms_per_batch": e2e_fly_ms},
                # same API, dataset resident on the device as CSR (reference data/cupy.py), tiles densified per step
                "csr_resident_dataset": {"value": csr_value, "unit": "cells/s", "h2d_bytes_per_step": 0,
                                         "nnz_per_step": csr_nnz}},
        "with_optimizer": with_opt,
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "loss_last": loss_last, "status": status,
        "all_reduce": all_reduce, "dp_check": dp_check, "other_configs": other_configs,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--batches", type=int, default=20, help="distinct resident minibatches")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels directly instead of replaying CUDA graphs")
    ap.add_argument("--other-configs", default=None,
                    help="comma-separated BASELINE.json configs measured briefly into `other_configs` (default: 2 on one "
                         "GPU, 4 and 5 on eight; '' = none)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_native(args, cfg)


if __name__ == "__main__":
    main()
