"""Host-side runtime of one model: flat parameter / gradient / BatchNorm buffers, workspace
and the ctypes calls into the C ABI.  PyTorch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"drvish_b200: {what} must be a CUDA tensor (there is no CPU fallback)")


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    _require_cuda(t, what)
    if t.dtype != torch.float32:
        raise RuntimeError(f"drvish_b200: {what} must be float32 (the reference computes in fp32), got {t.dtype}")
    return t.contiguous()


def make_dims(n_cells, n_genes, n_latent, layers, n_drugs=0, n_classes=0, n_doses_total=0) -> _lib.Dims:
    if len(layers) < 1 or len(layers) > _lib.DRV_MAX_LAYERS:
        raise ValueError(f"between 1 and {_lib.DRV_MAX_LAYERS} hidden layers are supported, got {len(layers)}")
    d = _lib.Dims()
    d.n_cells, d.n_genes, d.n_latent, d.n_layers = int(n_cells), int(n_genes), int(n_latent), len(layers)
    for i, h in enumerate(layers):
        d.layers[i] = int(h)
    d.n_drugs, d.n_classes, d.n_doses_total = int(n_drugs), int(n_classes), int(n_doses_total)
    return d


class ModelRuntime:
    """Owns the device-side state the kernels read and write for one NBVAE / DRNBVAE."""

    def __init__(self, model):
        self.model = model
        self.lib = _lib.load()
        self.device: Optional[torch.device] = None
        self.flat_params = self.flat_grads = self.flat_bn = self.bn_counters = None
        self._param_views: List[torch.Tensor] = []
        self._grad_views: List[torch.Tensor] = []
        self._params: List[nn.Parameter] = []
        self._workspaces: Dict[int, torch.Tensor] = {}
        self._dims: Dict[int, _lib.Dims] = {}
        self._out = None
        self._out_host = None
        self._tail_host = None
        self.last_launches = 0
        self.last_path = 0
        self._scalars: Dict[int, tuple] = {}
        self._hp_cache: Dict[tuple, _lib.HParams] = {}
        m = model
        self.n_genes = m.encoder.fc[0].num_features
        self.layers = [m.encoder.fc[i].out_features for i in range(2, len(m.encoder.fc) - 1, 3)]
        self.n_latent = m.n_latent
        self.lmbs = list(getattr(m, "lmb_list", []))
        self.n_drugs = len(self.lmbs)
        self.n_classes = int(getattr(m, "n_classes", 0) or 0)
        self.doses = [l.n_conditions for l in self.lmbs]
        self.dose_offsets_host = [0]
        for c in self.doses:
            self.dose_offsets_host.append(self.dose_offsets_host[-1] + c)
        self.n_doses_total = self.dose_offsets_host[-1]
        self.dose_offsets = None

    # ------------------------------------------------------------------------------------
    def dims(self, n_cells: int) -> _lib.Dims:
        d = self._dims.get(n_cells)
        if d is None:
            d = make_dims(n_cells, self.n_genes, self.n_latent, self.layers, self.n_drugs, self.n_classes,
                          self.n_doses_total)
            self._dims[n_cells] = d
        return d

    def _scalar(self, v) -> float:
        """Python float of a scalar hyper-parameter.  With the reference's ``device=`` constructor argument
        ``epsilon`` / ``l_loc`` / ``l_scale`` / ``sigma_scale`` are CUDA tensors (nbvae.py:44-47): ``float()`` on them
        is a blocking device-to-host copy - and illegal during CUDA-graph capture - so each tensor is read ONCE and
        remembered by (identity, version counter)."""
        if not torch.is_tensor(v):
            return float(v)
        key = (id(v), v._version)
        hit = self._scalars.get(id(v))
        if hit is None or hit[0] != key:
            hit = (key, float(v.reshape(-1)[0]), v)  # the tensor is kept alive so that its id stays unique
            self._scalars[id(v)] = hit
        return hit[1]

    def hparams(self, flags: int = 0, include_guide_factor: bool = True) -> _lib.HParams:
        m = self.model
        flags = int(flags) | int(os.environ.get("DRVISH_B200_DEBUG_FLAGS", "0"))
        lmb = self.lmbs[0] if self.lmbs else None
        bn0 = m.encoder.fc[0]
        vals = (self._scalar(m.scale_factor), self._scalar(m.epsilon), self._scalar(m.l_loc), self._scalar(m.l_scale),
                float(lmb.lam_scale) if lmb else 1.0, float(lmb.prior_bias_scale) if lmb else 1.0,
                self._scalar(getattr(m, "sigma_scale", 1.0)), float(lmb.alpha) if lmb else 1.0, float(bn0.eps),
                float(bn0.momentum), 1 if include_guide_factor else 0, flags)
        hp = self._hp_cache.get(vals)
        if hp is None:
            hp = _lib.HParams()
            (hp.scale_factor, hp.epsilon, hp.lib_loc, hp.lib_scale, hp.lam_scale, hp.bias_scale, hp.sigma_scale, hp.alpha,
             hp.bn_eps, hp.bn_momentum, hp.include_guide_factor, hp.flags) = vals
            if len(self._hp_cache) > 64:
                self._hp_cache.clear()
            self._hp_cache[vals] = hp
        return hp

    # ------------------------------------------------------------------------------------
    def _ordered_params(self):
        """(tensor-group list) in the order of drv_param_offsets."""
        m = self.model
        groups = []
        for stack in (m.encoder.fc, m.decoder.fc):
            for i in range(0, len(stack), 3):
                bn, lin = stack[i], stack[i + 2]
                groups += [[bn.weight], [bn.bias], [lin.weight], [lin.bias]]
        if self.lmbs:
            groups.append([l.weight_loc for l in self.lmbs])
            groups.append([l.weight_scale_unconstrained for l in self.lmbs])
            groups.append([l.bias_loc for l in self.lmbs])
            groups.append([l.bias_scale_unconstrained for l in self.lmbs])
        return groups

    def _bn_modules(self):
        m = self.model
        return [stack[i] for stack in (m.encoder.fc, m.decoder.fc) for i in range(0, len(stack), 3)]

    def ensure_flat(self) -> None:
        """(Re)build the flat buffers if parameters moved (``.cuda()``, ``.to()``) since the last call."""
        first = self.model.encoder.fc[0].weight
        _require_cuda(first, "model parameters")
        if (self.flat_params is not None and self._params and self._params[0].data_ptr() == self._param_views[0].data_ptr()
                and self._params[-1].data_ptr() == self._param_views[-1].data_ptr()):
            return
        dev = first.device
        self.device = dev
        d = self.dims(2)
        groups = self._ordered_params()
        n = len(groups)
        offs = (C.c_int64 * n)()
        sizes = (C.c_int64 * n)()
        total = C.c_int64()
        got = self.lib.drv_param_offsets(C.byref(d), offs, sizes, n, C.byref(total))
        if got != n:
            raise RuntimeError(f"drvish_b200: parameter layout mismatch ({got} tensors from the library, {n} here)")
        flat = torch.zeros(total.value, dtype=torch.float32, device=dev)
        # 32 spare floats: the first DRV_GRAD_TAIL_FLOATS carry the loss and the status bits under data parallel
        gflat = torch.zeros(total.value + 32, dtype=torch.float32, device=dev)
        self._params, self._param_views, self._grad_views = [], [], []
        for gi, group in enumerate(groups):
            off = offs[gi]
            if sum(p.numel() for p in group) != sizes[gi]:
                raise RuntimeError("drvish_b200: parameter size mismatch with the library layout")
            for p in group:
                k = p.numel()
                view = flat[off:off + k].view(p.shape)
                view.copy_(p.data.to(dev, torch.float32))
                p.data = view
                self._params.append(p)
                self._param_views.append(view)
                self._grad_views.append(gflat[off:off + k].view(p.shape))
                off += k
        self.flat_params, self.flat_grads = flat, gflat
        self.n_param_floats = total.value
        # float offset of the first decoder tensor: [decoder_offset:] is the slice whose gradients are
        # final before the encoder backward starts (decoder + dose-response head + loss slot)
        self.decoder_offset = int(offs[4 * (len(self.layers) + 1)])
        # BatchNorm buffers
        bns = self._bn_modules()
        nb = 2 * len(bns)
        boffs = (C.c_int64 * nb)()
        bsizes = (C.c_int64 * nb)()
        btotal = C.c_int64()
        got = self.lib.drv_bn_offsets(C.byref(d), boffs, bsizes, nb, C.byref(btotal))
        if got != nb:
            raise RuntimeError("drvish_b200: BatchNorm layout mismatch")
        bflat = torch.zeros(btotal.value, dtype=torch.float32, device=dev)
        counters = torch.zeros(len(bns), dtype=torch.int64, device=dev)
        for i, bn in enumerate(bns):
            for j, name in enumerate(("running_mean", "running_var")):
                off, k = boffs[2 * i + j], bsizes[2 * i + j]
                view = bflat[off:off + k]
                view.copy_(getattr(bn, name).to(dev, torch.float32))
                bn._buffers[name] = view
            counters[i] = bn.num_batches_tracked.to(dev)
            bn._buffers["num_batches_tracked"] = counters[i]
        self.flat_bn, self.bn_counters = bflat, counters
        if self.n_drugs:
            self.dose_offsets = torch.tensor(self.dose_offsets_host, dtype=torch.int32, device=dev)
        self._workspaces.clear()
        self._out = torch.zeros(2, dtype=torch.float64, device=dev)
        self._out_host = torch.zeros(2, dtype=torch.float64).pin_memory()

    def workspace(self, n_cells: int) -> torch.Tensor:
        ws = self._workspaces.get(n_cells)
        if ws is None:
            nbytes = self.lib.drv_workspace_bytes(C.byref(self.dims(n_cells)))
            if nbytes == 0:
                raise RuntimeError("drvish_b200: invalid problem shape (need >= 2 cells per minibatch)")
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._workspaces[n_cells] = ws
        return ws

    def shape_padded(self, n_cells: int) -> bool:
        """Does drv_elbo_step run this shape as a padded (aligned) problem (k_pad.cu)?"""
        return bool(self.lib.drv_shape_padded(C.byref(self.dims(n_cells))))

    def rebase_grads(self, new_flat: torch.Tensor) -> None:
        """Move the flat gradient buffer (and every ``param.grad`` view) onto ``new_flat`` - e.g. the peer-mapped
        buffer of ``parallel.PeerAllReduce``.  Same layout; graphs captured earlier must be dropped by the caller."""
        if new_flat.numel() < self.flat_grads.numel() or new_flat.dtype != torch.float32 or new_flat.device != self.flat_grads.device:
            raise ValueError("rebase_grads: incompatible buffer")
        new_flat[:self.flat_grads.numel()].copy_(self.flat_grads)
        base = self.flat_grads.data_ptr()
        self._grad_views = [new_flat[(v.data_ptr() - base) // 4:(v.data_ptr() - base) // 4 + v.numel()].view(v.shape)
                            for v in self._grad_views]
        self.flat_grads = new_flat

    def attach_grads(self) -> None:
        """Point every ``param.grad`` at its slice of the flat gradient buffer (what Pyro's
        ``loss.backward()`` leaves behind)."""
        for p, g in zip(self._params, self._grad_views):
            p.grad = g

    # ------------------------------------------------------------------------------------
    def elbo_step(self, x, labels, y_flat, noise: dict, want_grads: bool, flags: int = 0,
                  include_guide_factor: bool = True, mcv=None):
        """Enqueue one drv_elbo_step (or drv_elbo_step_mcv when ``mcv = (x1, log_split, log_split_complement)``,
        each of the last two a contiguous fp32 [B] device tensor).  Returns nothing; read the result with read_out()."""
        self.ensure_flat()
        x = _f32c(x, "x")
        if x.dim() != 2 or x.shape[1] != self.n_genes:
            raise RuntimeError(f"drvish_b200: x must be (cells, {self.n_genes}), got {tuple(x.shape)}")
        B = x.shape[0]
        d = self.dims(B)
        hp = self.hparams(flags, include_guide_factor)
        ws = self.workspace(B)
        if self.n_drugs:
            if labels is None or y_flat is None:
                raise RuntimeError("drvish_b200: DRNBVAE step needs labels and y")
            _require_cuda(labels, "labels")
            if labels.dtype != torch.int64:
                labels = labels.long()
            labels = labels.contiguous()
        if mcv is not None:
            x1, ls, lsc = mcv
            x1 = _f32c(x1, "x1")
            if x1.shape != x.shape or ls.numel() != B or lsc.numel() != B:
                raise RuntimeError("drvish_b200: MCV step needs x1 of x0's shape and one data-split term per cell")
            code = self.lib.drv_elbo_step_mcv(
                C.byref(d), C.byref(hp), _ptr(x), _ptr(x1), _ptr(ls), _ptr(lsc), _ptr(noise["eps_z"]), _ptr(noise["eps_l"]),
                _ptr(self.flat_params), _ptr(self.flat_grads) if want_grads else None, _ptr(self.flat_bn),
                _ptr(self.bn_counters), _ptr(self._out), _ptr(ws), ws.numel(), _stream(self.device))
            _lib.check(code, "drv_elbo_step_mcv")
            self.last_launches = self.lib.drv_last_launch_count()
            return
        code = self.lib.drv_elbo_step(
            C.byref(d), C.byref(hp), _ptr(x), _ptr(labels) if self.n_drugs else None,
            _ptr(y_flat) if self.n_drugs else None, _ptr(noise["eps_z"]), _ptr(noise["eps_l"]),
            _ptr(noise.get("w_prior")), _ptr(noise.get("b_prior")), _ptr(noise.get("eps_bias")),
            _ptr(self.dose_offsets), _ptr(self.flat_params), _ptr(self.flat_grads) if want_grads else None,
            _ptr(self.flat_bn), _ptr(self.bn_counters), _ptr(self._out), _ptr(ws), ws.numel(), _stream(self.device))
        _lib.check(code, "drv_elbo_step")
        self.last_launches = self.lib.drv_last_launch_count()

    def read_tail(self):
        """Data parallel, after the all-reduce: one 16-byte D2H copy of the gradient buffer's tail slots ->
        (summed loss, status bits raised by ANY rank).  Identical on every rank."""
        if self._tail_host is None:
            self._tail_host = torch.zeros(_lib.DRV_GRAD_TAIL_FLOATS, dtype=torch.float32).pin_memory()
        n = self.n_param_floats
        self._tail_host.copy_(self.flat_grads[n:n + _lib.DRV_GRAD_TAIL_FLOATS], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        t = self._tail_host.tolist()
        status = ((_lib.DRV_STATUS_NAN_LOSS if t[1] > 0 else 0) | (_lib.DRV_STATUS_MISSING_CLASS if t[2] > 0 else 0)
                  | (_lib.DRV_STATUS_BAD_LABEL if t[3] > 0 else 0))
        return t[0], status

    def read_out(self):
        """One 16-byte D2H copy: (loss, status)."""
        self._out_host.copy_(self._out, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        loss = float(self._out_host[0])
        words = self._out_host[1:2].view(torch.int32)
        status = int(words[0])
        self.last_path = int(words[1])  # drv_step_out.path: bit 0 tcgen05 decoder, bit 1 tcgen05 encoder input layer
        return loss, status

    # ------------------------------------------------------------------------------------
    def encoder_forward(self, x):
        self.ensure_flat()
        x = _f32c(x, "x")
        B, L = x.shape[0], self.n_latent
        d, hp, ws = self.dims(B), self.hparams(), self.workspace(B)
        z_loc = torch.empty(B, L, device=self.device)
        z_scale = torch.empty(B, L, device=self.device)
        l_loc = torch.empty(B, 1, device=self.device)
        l_scale = torch.empty(B, 1, device=self.device)
        code = self.lib.drv_encoder_forward(C.byref(d), C.byref(hp), _ptr(x), _ptr(self.flat_params), _ptr(self.flat_bn),
                                            _ptr(self.bn_counters), _ptr(z_loc), _ptr(z_scale), _ptr(l_loc),
                                            _ptr(l_scale), _ptr(ws), ws.numel(), _stream(self.device))
        _lib.check(code, "drv_encoder_forward")
        return z_loc, z_scale, l_loc, l_scale

    def decoder_forward(self, z, library):
        self.ensure_flat()
        z = _f32c(z, "z")
        B, G = z.shape[0], self.n_genes
        library = _f32c(library.expand(B, 1) if library.dim() == 2 else library.reshape(-1, 1).expand(B, 1), "library")
        d, hp, ws = self.dims(B), self.hparams(), self.workspace(B)
        log_r = torch.empty(B, G, device=self.device)
        logit = torch.empty(B, G, device=self.device)
        code = self.lib.drv_decoder_forward(C.byref(d), C.byref(hp), _ptr(z), _ptr(library), _ptr(self.flat_params),
                                            _ptr(self.flat_bn), _ptr(self.bn_counters), _ptr(log_r), _ptr(logit),
                                            _ptr(ws), ws.numel(), _stream(self.device))
        _lib.check(code, "drv_decoder_forward")
        return log_r, logit


def logit_mean_sigmoid(z: torch.Tensor, weights: Sequence[torch.Tensor], biases: Sequence[torch.Tensor],
                       labels: torch.Tensor, n_classes: Optional[int] = None) -> List[torch.Tensor]:
    """All drugs at once: weights[i] (L,1), biases[i] (C_i,1) -> list of (K, C_i, 1).
    Raises like the reference when a class has no cell in the batch (modules.py:175-179)."""
    lib = _lib.load()
    z = _f32c(z, "z")
    dev = z.device
    B, L = z.shape
    D = len(weights)
    doses = [int(b.shape[0]) for b in biases]
    offs = [0]
    for c in doses:
        offs.append(offs[-1] + c)
    labels = labels.to(dev).long().contiguous()
    K = int(n_classes) if n_classes else int(labels.max().item()) + 1
    w = torch.cat([_f32c(t.detach(), "weights").reshape(1, L) for t in weights], 0).contiguous()
    b = torch.cat([_f32c(t.detach(), "bias").reshape(-1) for t in biases], 0).contiguous()
    d = make_dims(B, 1, L, [1], D, K, offs[-1])
    out = torch.empty(K * offs[-1], device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(4 * B * D, dtype=torch.uint8, device=dev)
    doff = torch.tensor(offs, dtype=torch.int32, device=dev)
    code = lib.drv_logit_mean_sigmoid(C.byref(d), _ptr(z), _ptr(w), _ptr(b), _ptr(labels), _ptr(doff), _ptr(out),
                                      _ptr(status), _ptr(ws), ws.numel(), _stream(dev))
    _lib.check(code, "drv_logit_mean_sigmoid")
    st = int(status.item())
    if st & (_lib.DRV_STATUS_MISSING_CLASS | _lib.DRV_STATUS_BAD_LABEL):
        raise RuntimeError("index out of bounds: every class 0..K-1 must occur in the batch "
                           "(reference modules.py:175-179 raises the same way)")
    return [out[K * offs[i]:K * offs[i + 1]].view(K, doses[i], 1) for i in range(D)]
