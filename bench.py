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


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = min(args.steps, 20)
    rate, ms, cores, B = oracle_cpu_rate(cfg, steps, min(args.warmup, 3))
    sample = f"{steps} steps of one {B}-cell minibatch of {cfg['name']}"
    line = {
        "impl": "reference", "metric": "cells/sec per SVI step (ELBO fwd+bwd)", "value": rate, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": min(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["name"], "minibatch_per_step": B, "device": "host CPU"},
        "cpu_baseline": {"value": rate, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_native(args, cfg):
    import torch.distributed as dist

    import drvish_b200 as D
    from drvish_b200 import _lib
    from drvish_b200.parallel import allreduce_grads_and_loss

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
    model = build_native_model(cfg, device)
    rt = model._runtime()
    elbo = D.FusedTraceELBO(seed=1234, data_parallel=(world > 1), use_cuda_graph=not args.no_graph)
    n_batches = max(20, args.batches)
    xs, labels, y = synth_batches(cfg, n_batches, 1234 + 3 + rank, device)
    B, G = cfg["batch"], cfg["n_input"]

    def batch_args(i, x=None):
        a = [xs[i % n_batches] if x is None else x]
        if cfg["drugs"]:
            a += [labels, y]
        return a

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (`value`) ----------------------------------------------------
    # same code path as loss_and_grads minus the per-step loss read-back (one sync at the end)
    yflat = elbo._flatten_y(rt, y, device) if cfg["drugs"] else None
    rt.ensure_flat()
    if world > 1 and not elbo.dp_overlap:
        elbo.dp_prepare(rt)  # peer-mapped gradient buffer (before any graph is captured)

    def device_step(i, single_call=False):
        # exactly what FusedTraceELBO.loss_and_grads enqueues, minus the per-step loss read-back
        xb = xs[i % n_batches]
        if world > 1 and not single_call and not elbo.dp_overlap:
            elbo._enqueue(rt, xb, labels, yflat, True, part=3)
            elbo.dp_all_reduce(rt)
        elif world > 1 and not single_call:
            elbo._enqueue(rt, xb, labels, yflat, True, part=1)
            w1 = dist.all_reduce(rt.flat_grads[rt.decoder_offset:], op=dist.ReduceOp.SUM, async_op=True)
            elbo._enqueue(rt, xb, labels, yflat, True, part=2)
            w2 = dist.all_reduce(rt.flat_grads[:rt.decoder_offset], op=dist.ReduceOp.SUM, async_op=True)
            w1.wait()
            w2.wait()
        else:
            elbo._enqueue(rt, xb, labels, yflat, True)
        elbo.step_index += 1

    if not args.no_graph:
        for i in range(2 * n_batches):  # untimed: every resident minibatch is seen twice -> its graph is captured
            device_step(i)
    for i in range(args.warmup):
        device_step(i)
    launches_per_step = rt.last_launches + 1  # + the noise draw (one launch)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        device_step(args.warmup + i)
    ev1.record()
    barrier()
    clocks = sampler.finish()
    ms_total = ev0.elapsed_time(ev1)
    loss_last, status = rt.read_out()
    t = torch.tensor([ms_total], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- the same loop with the fused optimizer step behind every ELBO step (AggMo, 3 betas,
    # nesterov, per-tensor clip_norm 20 as in the reference notebook); an extra, the headline excludes it
    opt = D.FusedAggMo(model, lr=5e-4, betas=[0.0, 0.9, 0.99], nesterov=True, clip_norm=20.0)
    params_before = rt.flat_params.clone()
    for i in range(3):
        device_step(i)
        opt.step()
    barrier()
    ev0.record()
    for i in range(args.steps):
        device_step(args.warmup + i)
        opt.step()
    ev1.record()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_opt = float(t.item()) / args.steps
    with_opt = {"value": world * B / (ms_opt * 1e-3), "unit": "cells/s", "ms_per_step": ms_opt,
                "optimizer": "FusedAggMo(betas=[0,0.9,0.99], nesterov, clip_norm=20): drv_aggmo_step, 3-4 launches",
                "optimizer_ms": ms_opt - ms_per_step}
    rt.flat_params.copy_(params_before)  # the measurements below run on the initial weights again

    # ---- live kernel timing for the roofline: CUDA events recorded by the library on the step's
    # own stream, around the sections and around the single kernels of the fused decoder path.
    # The side stream is switched off here (flag 64) so that a section contains ALL of its work.
    n_prof = min(args.steps, 32)
    os.environ["DRVISH_B200_DEBUG_FLAGS"] = "64"
    elbo.use_cuda_graph = False  # direct launches: the profiler's events are recorded by the library
    for i in range(3):
        device_step(i, single_call=True)
    torch.cuda.synchronize()
    lib.drv_profile_enable(1)
    for i in range(n_prof):
        device_step(i, single_call=True)
    torch.cuda.synchronize()
    sec = (C.c_float * 7)()
    kms = (C.c_float * 5)()
    nprof = C.c_int()
    lib.drv_profile_read(sec, C.byref(nprof))
    lib.drv_profile_read_kernels(kms)
    lib.drv_profile_enable(0)
    os.environ["DRVISH_B200_DEBUG_FLAGS"] = "0"
    elbo.use_cuda_graph = not args.no_graph

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

    def run_e2e(pf, dataset, n):
        pf.batches = (dataset[i % len(dataset)] for i in range(n))
        last = None
        for batch in pf:
            last = elbo.loss_and_grads(model.model, model.guide, *batch)
        return last

    def time_e2e(dataset, on_the_fly=False):
        pf = PinnedBatchPrefetcher(None, device, pack_on_the_fly=on_the_fly)
        run_e2e(pf, dataset, 3 * pf.SLOTS)  # untimed: every staging slot gets its CUDA graph
        barrier()
        pf.h2d_bytes = 0
        t0 = time.perf_counter()
        run_e2e(pf, dataset, e2e_steps)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * B * e2e_steps / float(t.item()), pf.h2d_bytes // e2e_steps

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
    t = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    csr_value = world * B * e2e_steps / float(t.item())
    csr_nnz = int(csr.data.numel() // n_csr)

    e2e_value, e2e_h2d = time_e2e(fp32_set)
    e2e_u16, e2e_u16_h2d = time_e2e(u16_set)
    e2e_fly, e2e_fly_h2d = time_e2e(fp32_set, on_the_fly=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the fused decoder-last-layer + NB-likelihood step ----------------
    peaks = load_peaks()
    Bf, Gf, hf = float(B), float(G), float(cfg["layers"][0])
    bytes_alg, flops_alg = decoder_algorithmic(cfg)
    t_dec_ms = sec[2] + sec[4]
    t_hbm = bytes_alg / (peaks["hbm"] * 1e9)
    t_tensor = flops_alg / (peaks["tensor"] * 1e12)
    if t_tensor >= t_hbm:
        roof = {"bound": "tensor", "achieved": flops_alg / (t_dec_ms * 1e-3) / 1e12, "peak": peaks["tensor"],
                "unit": "TFLOP/s"}
    else:
        roof = {"bound": "hbm", "achieved": bytes_alg / (t_dec_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"cfg{args.config}", {})
    roof["traffic"] = traffic.get("decoder_step_bytes")
    roof["kernel"] = ("fused decoder step = k_dec<1> (logsumexp) + k_dec<2> (NB likelihood + NB part of dlogits) + "
                      "k_dec<3> (softmax coupling) + dAct and dW tcgen05 GEMMs (sections DEC_NB_FWD + DEC_NB_BWD, "
                      "side stream off)")
    roof["algorithmic"] = {"bytes": bytes_alg, "flops": flops_alg, "formula": "8BG+8Bh+16Gh bytes, 12BGh flops (SURVEY 8d)"}
    roof["peak_source"] = peaks["which"] + " (bf16 sustained cuBLAS; the decoder kernels use kind::f16 - fp16 operands, fp32 accumulate - whose nominal peak is the bf16 one)"
    roof["section_ms"] = {k: round(float(sec[i]), 4) for i, k in enumerate(
        ["enc_fwd", "dec_trunk_fwd", "dec_nb_fwd", "dose_response", "dec_nb_bwd", "dec_trunk_bwd", "enc_bwd"])}
    roof["profiled_steps"] = int(nprof.value)
    # per-kernel entries: algorithmic work of that launch / its own event-timed duration
    kdefs = [
        ("k_dec<1> logsumexp", 4 * Gf * hf + 4 * Bf * hf, 2 * Bf * Gf * hf),
        ("k_dec<2> NB likelihood (+ NB part of dlogits)", 4 * Bf * Gf + 8 * Bf * Gf + 8 * Gf * hf + 4 * Bf * hf, 4 * Bf * Gf * hf),
        ("k_dec<3> softmax coupling", 8 * Bf * Gf + 4 * Gf * hf + 4 * Bf * hf, 2 * Bf * Gf * hf),
        ("dAct GEMM", 8 * Bf * Gf + 8 * Gf * hf + 4 * Bf * hf, 4 * Bf * Gf * hf),
        ("dW GEMM", 8 * Bf * Gf + 8 * Gf * hf + 4 * Bf * hf, 4 * Bf * Gf * hf),
    ]
    rk = []
    for i, (name, kb, kf) in enumerate(kdefs):
        ms = float(kms[i])
        if ms <= 0:
            continue
        th, tt = kb / (peaks["hbm"] * 1e9), kf / (peaks["tensor"] * 1e12)
        ent = {"kernel": name, "ms": round(ms, 4), "bound": "hbm" if th >= tt else "tensor"}
        if th >= tt:
            ent.update(achieved=kb / (ms * 1e-3) / 1e9, peak=peaks["hbm"], unit="GB/s")
        else:
            ent.update(achieved=kf / (ms * 1e-3) / 1e12, peak=peaks["tensor"], unit="TFLOP/s")
        ent["frac"] = ent["achieved"] / ent["peak"]
        ent["traffic"] = traffic.get(name.split(" ")[0])
        rk.append(ent)
    roof["kernels"] = rk

    # ---- CPU baseline on the host cores (bounded sample) ---------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        rate, ms, cores, Bc = oracle_cpu_rate(cfg, args.cpu_steps, 2)
        cpu = {"value": rate, "unit": "cells/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} timed steps (2 warm-up) of one {Bc}-cell minibatch, fp32, PyTorch-eager "
                         "restatement of the Pyro step (oracle/drvish_oracle.py)", "ms_per_step": ms}

    line = {
        "metric": "cells/sec per SVI step (ELBO fwd+bwd)", "value": value, "unit": "cells/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 / tf32 tensor-core operands with fp32 accumulation where stated in DESIGN.md)",
        "data": "synthetic",
        "config": {"workload": cfg["name"], "minibatch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "l2": f"inputs rotate over {n_batches} resident minibatches "
                   f"({n_batches * B * G * 4 / 1e9:.2f} GB > 126 MB L2), no flush", "optimizer": "excluded from value/e2e (see with_optimizer)",
                   "noise": "Philox in-kernel draws each step", "cuda_graph": not args.no_graph,
                   "all_reduce": ("none (1 GPU)" if world == 1 else
                                  "two-part step, NCCL all-reduces overlapping the encoder backward" if elbo.dp_overlap else
                                  "one drv_peer_allreduce kernel over NVLink peer memory per step" if elbo._peer.get(id(rt)) is not None
                                  else "one NCCL all-reduce per step")},
        "e2e": {"value": e2e_value, "unit": "cells/s", "h2d_bytes_per_step": e2e_h2d, "d2h_bytes_per_step": 16,
                "steps": e2e_steps, "api": "FusedTraceELBO.loss_and_grads(model.model, model.guide, *batch) fed by PinnedBatchPrefetcher "
                       "(pinned host fp32 minibatches, H2D of the batches ahead overlapped with the running step)",
                "bound": "PCIe: %.1f GB/s host->device" % (e2e_h2d * e2e_value / (world * B) / 1e9),
                # same API, minibatches packed ONCE to lossless uint16 when the dataset is built
                "packed_u16_dataset": {"value": e2e_u16, "unit": "cells/s", "h2d_bytes_per_step": e2e_u16_h2d},
                # same API, fp32 host minibatches packed by the prefetcher's worker thread every step
                "packed_u16_on_the_fly": {"value": e2e_fly, "unit": "cells/s", "h2d_bytes_per_step": e2e_fly_h2d},
                # same API, dataset resident on the device as CSR (reference data/cupy.py), tiles densified per step
                "csr_resident_dataset": {"value": csr_value, "unit": "cells/s", "h2d_bytes_per_step": 0,
                                         "nnz_per_step": csr_nnz}},
        "with_optimizer": with_opt,
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "loss_last": loss_last, "status": status,
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_native(args, cfg)


if __name__ == "__main__":
    main()
