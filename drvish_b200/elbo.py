"""The fused ELBO: a drop-in for ``pyro.infer.Trace_ELBO`` / ``TraceGraph_ELBO`` as the ``loss=``
of ``pyro.infer.SVI`` for the drvish models (reference notebook cell 10; call sites
src/drvish/train/__init__.py:46,90-93).

``SVI.step(*batch)`` calls ``loss.loss_and_grads(model, guide, *batch)`` and
``SVI.evaluate_loss`` calls ``loss.loss(model, guide, *batch)``; both are implemented here by ONE
call into ``drv_elbo_step`` (no tracing, no autograd graph).  Gradients are left in
``param.grad`` exactly where Pyro's ``loss.backward()`` would leave them, BatchNorm running
buffers are updated in place, and the Python float loss is returned.

Data parallel: with ``torch.distributed`` initialised (one process per GPU) the
flat gradient buffer is summed over ranks with ONE all-reduce per step; the loss and the status
bits ride in the tail slots of the same buffer, and every rank RETURNS the summed loss and raises
on the status bits of ANY rank (so that control flow that depends on the loss - the plateau test of
``train_until_plateau``, train/__init__.py:97-106 - and errors stay in step on all ranks).
``loss()`` (``SVI.evaluate_loss``) all-reduces its (loss, status) record the same way.  The R-rank
result is sum_r reference(shard_r).
"""
from __future__ import annotations

import os
import warnings
from typing import Optional

import torch

from . import _lib
from .models import HAVE_PYRO, NBVAE
from .parallel import PeerAllReduce, allreduce_grads_and_loss

if HAVE_PYRO:  # pragma: no cover
    import pyro


class FusedTraceELBO:
    """:param seed: seed of the counter-based (Philox) noise generator used for the
        reparameterised draws; every step consumes a fresh counter range.
    :param data_parallel: ``None`` = all-reduce gradients iff ``torch.distributed`` is initialised.
    :param include_guide_factor: keep the guide-side monotonic-bias factor (pathwise gradient).
    :param force_simt: never use the tcgen05 kernels (A/B parity checks).
    """

    STAGING_AFTER_MISSES = 64  # first sightings in a row (per shape) before batches are copied into a staging tile

    def __init__(self, num_particles: int = 1, seed: int = 0, data_parallel: Optional[bool] = None,
                 process_group=None, include_guide_factor: bool = True, force_simt: bool = False,
                 update_bn_stats: bool = True, extra_flags: int = 0, use_cuda_graph: bool = True,
                 **_ignored):
        if num_particles != 1:
            raise NotImplementedError("the fused ELBO evaluates one particle (the reference notebook's setting)")
        self.seed = int(seed)
        self.step_index = 0
        self.data_parallel = data_parallel
        self.process_group = process_group
        self.include_guide_factor = include_guide_factor
        self.flags = ((_lib.DRV_FLAG_FORCE_SIMT if force_simt else 0) | (0 if update_bn_stats else _lib.DRV_FLAG_NO_BN_UPDATE)
                      | int(extra_flags))
        self._explicit_noise = None
        self._noise_bufs = {}
        self.last_status = 0
        self.use_cuda_graph = bool(use_cuda_graph)
        # data parallel: True = two-part step, the all-reduce of the decoder slice overlaps the encoder backward;
        # False = one graph + one all-reduce after it (see DESIGN.md section 5 for the measurements)
        self.dp_overlap = os.environ.get("DRV_DP_OVERLAP", "0") == "1"
        # the collective inside the step, decoder slice overlapped with the encoder backward (DRV_DP_FUSED=0: one
        # all-reduce kernel behind the step)
        self.dp_fused = os.environ.get("DRV_DP_FUSED", "1") == "1"
        self._graphs = {}
        self._graph_seen = {}
        self._staging = {}  # (shape, mode) -> fixed staging tile for loaders that never repeat a buffer address
        self._peer = {}  # id(runtime) -> PeerAllReduce or None (NCCL)
        self._step_ctr = None  # device uint64: steps taken (drives the Philox counters)

    # -- noise ------------------------------------------------------------------------------
    def set_noise(self, noise: Optional[dict]) -> None:
        """Parity mode: use these noise tensors (``eps_z`` (B,L), ``eps_l`` (B,1) and, for DRNBVAE,
        lists ``w_prior`` (L,1), ``b_prior`` (C_i,1), ``eps_bias`` (C_i,1)) instead of drawing."""
        self._explicit_noise = noise

    def _noise(self, rt, B: int, draw: bool = True) -> dict:
        dev = rt.device
        if self._explicit_noise is not None:
            n = self._explicit_noise
            out = {"eps_z": n["eps_z"].to(dev, torch.float32).contiguous(),
                   "eps_l": n["eps_l"].to(dev, torch.float32).reshape(-1).contiguous()}
            if rt.n_drugs:
                cat = lambda ts: torch.cat([t.to(dev, torch.float32).reshape(-1) for t in ts]).contiguous()
                out.update(w_prior=cat(n["w_prior"]), b_prior=cat(n["b_prior"]), eps_bias=cat(n["eps_bias"]))
            return out
        key = (B, str(dev))
        bufs = self._noise_bufs.get(key)
        L, D, T = rt.n_latent, rt.n_drugs, rt.n_doses_total
        if bufs is None:
            bufs = {"eps_z": torch.empty(B * L, device=dev), "eps_l": torch.empty(B, device=dev)}
            if D:
                bufs.update(w_prior=torch.empty(D * L, device=dev), b_prior=torch.empty(T, device=dev),
                            eps_bias=torch.empty(T, device=dev))
            self._noise_bufs[key] = bufs
        if not draw:
            return bufs  # second part of a two-part step: same draws
        # production mode: Philox counters keyed by (seed, step, rank); the step index is read from a
        # DEVICE counter so that a captured CUDA graph draws fresh noise on every replay.  The draw
        # order follows the reference (latent, library, then per-drug guide / prior draws - SURVEY 8b)
        import ctypes as C
        from .runtime import _ptr, _stream

        lib, st = rt.lib, _stream(dev)
        if self._step_ctr is None or self._step_ctr.device != dev:
            # {step counter, launch ticket}; the draw kernel advances the counter itself
            self._step_ctr = torch.tensor([self.step_index, 0], dtype=torch.int64, device=dev)
        per_step = 4 * ((B * L + 3) // 4 + (B + 3) // 4 + (D * L + 3) // 4 + 2 * ((T + 3) // 4))
        rank = torch.distributed.get_rank() if self._dp_enabled() else 0
        dims, hp = rt.dims(B), rt.hparams()
        _lib.check(lib.drv_draw_step_noise(C.byref(dims), C.byref(hp), C.c_uint64(self.seed), C.c_uint64(rank * per_step),
                                           _ptr(self._step_ctr), C.c_uint64(4096 * per_step), _ptr(bufs["eps_z"]),
                                           _ptr(bufs["eps_l"]), _ptr(bufs["w_prior"]) if D else None,
                                           _ptr(bufs["b_prior"]) if D else None, _ptr(bufs["eps_bias"]) if D else None,
                                           st), "drv_draw_step_noise")
        return bufs

    # -- helpers ----------------------------------------------------------------------------
    def dp_all_reduce(self, rt) -> None:
        """Sum the flat gradient buffer (+ loss slot) over the ranks: one peer-memory kernel over NVLink when the
        group is a single node of 2 / 4 / 8 GPUs (``parallel.PeerAllReduce``), else one NCCL all-reduce."""
        self.dp_prepare(rt)
        peer = self._peer[id(rt)]
        if peer is not None:
            peer.all_reduce()
        else:
            torch.distributed.all_reduce(rt.flat_grads[:rt.n_param_floats + _lib.DRV_GRAD_TAIL_FLOATS],
                                         op=torch.distributed.ReduceOp.SUM, group=self.process_group)

    def dp_step(self, rt, x, labels, y, mcv=None, reduce: bool = True) -> None:
        """One data-parallel step: the whole ELBO step, then ONE all-reduce of the flat gradient buffer.  With the
        peer-memory kernel the all-reduce is part of the step's CUDA graph (its flag barriers are numbered by a device
        epoch counter, so a replay is as good as a launch): no Python and no launch gap between the last gradient
        kernel and the collective.  With NCCL it is issued after the graph."""
        peer = self._peer.get(id(rt))
        if reduce and peer is not None and self.dp_fused and not rt.shape_padded(x.shape[0]):
            # the step launches the collective itself: the decoder slice is reduced on the library's side stream next
            # to the encoder backward, the encoder slice at the end (DRV_FLAG_FUSED_ALLREDUCE)
            self._enqueue(rt, x, labels, y, True, part=5, mcv=mcv)
            return
        in_graph = reduce and peer is not None and self.use_cuda_graph and self._explicit_noise is None
        self._enqueue(rt, x, labels, y, True, part=4 if in_graph else 3, mcv=mcv)
        if reduce and not in_graph:
            self.dp_all_reduce(rt)

    def dp_prepare(self, rt) -> None:
        """Set up the peer-memory all-reduce (collective call: every rank, before its first step)."""
        if id(rt) in self._peer:
            return
        peer = PeerAllReduce.create(rt.flat_grads.numel(), rt.device, self.process_group)
        if peer is not None:
            rt.rebase_grads(peer.data)
            peer.bind()  # drv_comm_bind: this thread's steps may launch the collective themselves
            self._graphs.clear()
            self._graph_seen.clear()
            self._staging.clear()
        self._peer[id(rt)] = peer

    def _dp_enabled(self) -> bool:
        if self.data_parallel is False:
            return False
        ok = torch.distributed.is_available() and torch.distributed.is_initialized()
        if self.data_parallel and not ok:
            raise RuntimeError("data_parallel=True needs torch.distributed.init_process_group first")
        return ok and torch.distributed.get_world_size(self.process_group) > 1

    @staticmethod
    def _module_of(model, guide) -> NBVAE:
        for fn in (model, guide):
            owner = getattr(fn, "__self__", None)
            if isinstance(owner, NBVAE):
                return owner
        if isinstance(model, NBVAE):
            return model
        raise TypeError("FusedTraceELBO expects the bound model/guide methods of a drvish_b200 NBVAE or DRNBVAE")

    def _flatten_y(self, rt, y, dev):
        if not rt.n_drugs:
            return None
        if torch.is_tensor(y):  # stacked (D, K, C, 1)
            y = list(y)
        # the bulk dose-response target is the same object every batch (reference data/target.py:31-34):
        # cache its flattened device copy (also keeps its address stable for CUDA-graph replay)
        sig = tuple((t.data_ptr(), t._version, tuple(t.shape)) for t in y)
        cached = getattr(self, "_y_cache", None)
        if cached is not None and cached[0] == sig:
            return cached[1]
        flat = self._flatten_y_uncached(rt, y, dev)
        self._y_cache = (sig, flat)
        return flat

    @staticmethod
    def _flatten_y_uncached(rt, y, dev):
        K = rt.n_classes
        parts = []
        for i, t in enumerate(y):
            if tuple(t.shape) != (K, rt.doses[i], 1):
                raise RuntimeError(f"y[{i}] must have shape ({K}, {rt.doses[i]}, 1) (reference drnbvae.py:112-116)")
            parts.append(t.to(dev, torch.float32).reshape(-1))
        return torch.cat(parts).contiguous()

    def _run(self, model, guide, args, want_grads: bool) -> float:
        module = self._module_of(model, guide)
        if HAVE_PYRO:  # pragma: no cover - registers the parameters so SVI's param capture sees them
            pyro.module(module.pyro_name, module)
        rt = module._runtime()
        rt.ensure_flat()
        x = args[0]
        mcv = None
        if getattr(module, "is_mcv", False):
            # (x0, x1, log_data_split, log_data_split_complement), mcv_nbvae.py:57-63
            if len(args) != 4:
                raise TypeError("MCVNBVAE batches are (x0, x1, log_data_split, log_data_split_complement)")
            B = x.shape[0]
            per_cell = lambda t: torch.as_tensor(t, dtype=torch.float32, device=x.device).reshape(-1).expand(B).contiguous() \
                if torch.as_tensor(t).numel() in (1, B) else None
            ls, lsc = per_cell(args[2]), per_cell(args[3])
            if ls is None or lsc is None:
                raise RuntimeError("drvish_b200: the data-split terms must be scalars or one value per cell")
            mcv = (args[1], ls, lsc)
            labels, y = None, None
        else:
            labels = args[1] if len(args) > 1 else None
            y = self._flatten_y(rt, args[2] if len(args) > 2 else None, rt.device)
        dp = self._dp_enabled()
        if want_grads and dp and not self.dp_overlap:
            # whole step, then ONE all-reduce of the flat gradient buffer (loss in its tail slot)
            if id(rt) not in self._peer:
                self.dp_prepare(rt)
            self.dp_step(rt, x, labels, y, mcv=mcv)
        elif want_grads and dp:
            # two-part step: the all-reduce of the decoder + dose-response slice (incl. the loss in the
            # tail slot) runs on NCCL's stream while the encoder backward is still computing
            self._enqueue(rt, x, labels, y, True, part=1, mcv=mcv)  # leaves the loss in flat_grads[n_param_floats]
            tail = rt.flat_grads[rt.decoder_offset:]
            w1 = torch.distributed.all_reduce(tail, op=torch.distributed.ReduceOp.SUM, group=self.process_group,
                                              async_op=True)
            self._enqueue(rt, x, labels, y, True, part=2, mcv=mcv)
            head = rt.flat_grads[:rt.decoder_offset]
            w2 = torch.distributed.all_reduce(head, op=torch.distributed.ReduceOp.SUM, group=self.process_group,
                                              async_op=True)
            w1.wait()
            w2.wait()
        else:
            self._enqueue(rt, x, labels, y, want_grads, mcv=mcv)
        self.step_index += 1
        if want_grads and dp:
            # the summed loss and the OR of all ranks' status bits, identical on every rank
            loss, status = rt.read_tail()
        else:
            loss, status = rt.read_out()
            if dp:  # evaluate_loss under data parallel: one small all-reduce of (loss, status bits)
                rec = torch.tensor([loss, float(status & 1), float((status >> 1) & 1), float((status >> 2) & 1)],
                                   dtype=torch.float64, device=rt.device)
                torch.distributed.all_reduce(rec, op=torch.distributed.ReduceOp.SUM, group=self.process_group)
                rec = rec.tolist()
                loss = rec[0]
                status = (1 if rec[1] > 0 else 0) | (2 if rec[2] > 0 else 0) | (4 if rec[3] > 0 else 0)
        self.last_status = status
        if status & (_lib.DRV_STATUS_MISSING_CLASS | _lib.DRV_STATUS_BAD_LABEL):
            raise RuntimeError("index out of bounds: every class 0..n_classes-1 must occur in the minibatch "
                               "(the reference's scatter_add_ raises the same way, modules.py:175-179)")
        if status & _lib.DRV_STATUS_NAN_LOSS:
            warnings.warn("Encountered NaN: loss")  # pyro.util.warn_if_nan
        if want_grads:
            rt.attach_grads()
        return loss

    def _enqueue(self, rt, x, labels, y, want_grads: bool, part: int = 0, mcv=None) -> None:
        """Noise draws + drv_elbo_step, either launched directly or replayed from a CUDA graph.

        A graph is keyed by the device pointers it bakes in (x, labels, y, mode, batch size); it is
        captured the second time a key is seen (the first call runs eagerly and warms up every
        kernel variant), so a training loop that cycles through a few staging buffers - e.g.
        PinnedBatchPrefetcher's two - replays ~70 kernel launches with one cudaGraphLaunch."""
        pflags = {0: 0, 1: _lib.DRV_FLAG_STOP_AFTER_DECODER, 2: _lib.DRV_FLAG_ENCODER_BACKWARD_ONLY,
                  3: _lib.DRV_FLAG_LOSS_IN_GRADS, 4: _lib.DRV_FLAG_LOSS_IN_GRADS,
                  5: _lib.DRV_FLAG_LOSS_IN_GRADS | _lib.DRV_FLAG_FUSED_ALLREDUCE}[part]

        def eager():
            # part 2 continues the step of part 1: same noise buffers, no new draws
            noise = self._noise(rt, x.shape[0], draw=(part != 2))
            rt.elbo_step(x, labels, y, noise, want_grads, self.flags | pflags, self.include_guide_factor, mcv=mcv)
            if part == 4:  # the step's collective, same stream: captured into the same graph
                self._peer[id(rt)].all_reduce()

        if not self.use_cuda_graph or self._explicit_noise is not None or not x.is_cuda:
            return eager()
        key = (x.data_ptr(), tuple(x.shape), 0 if labels is None else labels.data_ptr(), 0 if y is None else y.data_ptr(),
               want_grads, rt.flat_params.data_ptr(), part,
               None if mcv is None else tuple(t.data_ptr() for t in mcv))
        g = self._graphs.get(key)
        if g is not None:
            g.replay()
            return
        # A stock DataLoader hands over a FRESH tensor every step: no address ever repeats and no graph would ever be
        # captured.  Once that pattern shows (several first sightings in a row for one shape), the batch is copied into a
        # fixed staging tile per (shape, mode) - one device-to-device copy instead of ~60 direct launches - and the
        # graph of the staging tile is replayed.  Loaders that cycle through a few buffers (PinnedBatchPrefetcher,
        # resident datasets) keep their own addresses and never pay the copy.
        skey = (tuple(x.shape), want_grads, part, labels is not None, None if mcv is None else tuple(tuple(t.shape) for t in mcv))
        st = self._staging.get(skey)
        if st is not None and st["active"] and x.data_ptr() != st["x"].data_ptr():
            st["x"].copy_(x, non_blocking=True)
            if labels is not None:
                st["labels"].copy_(labels, non_blocking=True)
            smcv = None
            if mcv is not None:
                for dst, src in zip(st["mcv"], mcv):
                    dst.copy_(src, non_blocking=True)
                smcv = st["mcv"]
            return self._enqueue(rt, st["x"], st["labels"] if labels is not None else None, y, want_grads, part, smcv)
        seen = self._graph_seen.get(key, 0)
        if seen == 0 and part != 2:
            if st is None:
                st = self._staging[skey] = {"active": False, "misses": 0, "x": None, "labels": None, "mcv": None}
            st["misses"] += 1
            if st["misses"] > self.STAGING_AFTER_MISSES and not st["active"]:
                st["x"] = torch.empty_like(x)
                st["labels"] = torch.empty_like(labels) if labels is not None else None
                st["mcv"] = tuple(torch.empty_like(t) for t in mcv) if mcv is not None else None
                st["active"] = True
                return self._enqueue(rt, x, labels, y, want_grads, part, mcv)
        elif st is not None and not st["active"]:
            st["misses"] = 0  # an address came back: the loader recycles its buffers
        self._graph_seen[key] = seen + 1
        if seen < 1 or len(self._graphs) >= 64:
            if len(self._graph_seen) > 4096:
                self._graph_seen.clear()
            return eager()
        torch.cuda.current_stream(rt.device).synchronize()
        g = torch.cuda.CUDAGraph()
        try:
            # thread-local capture mode: a loader's worker thread (PinnedBatchPrefetcher) may allocate, query events
            # or synchronise its own stream while this thread captures - legal, but fatal under the default global mode
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                eager()
        except Exception as e:
            # capture unsupported here: stay on direct launches - and say so (the step is ~60 launches instead of 1)
            warnings.warn(f"drvish_b200: CUDA-graph capture of the ELBO step failed ({type(e).__name__}: {e}); "
                          "continuing with direct kernel launches")
            self.use_cuda_graph = False
            torch.cuda.synchronize()
            return eager()
        # the graph bakes in raw pointers only; whatever tensor later lives at the same address with
        # the same shape is what a replay reads, so no tensor needs to be kept alive here
        self._graphs[key] = g
        g.replay()

    # -- the two methods pyro.infer.SVI calls -----------------------------------------------------
    def loss(self, model, guide, *args, **kwargs) -> float:
        """``SVI.evaluate_loss``: forward only.  BatchNorm still normalises with batch statistics
        and advances its running buffers, as the reference does (it never calls ``.eval()``)."""
        return self._run(model, guide, args, want_grads=False)

    def loss_and_grads(self, model, guide, *args, **kwargs) -> float:
        """``SVI.step``: loss + every parameter gradient."""
        return self._run(model, guide, args, want_grads=True)

    def differentiable_loss(self, model, guide, *args, **kwargs):
        raise NotImplementedError("the fused ELBO has no autograd graph; use loss_and_grads")
