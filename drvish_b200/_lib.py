"""ctypes binding of ``libdrvish_b200.so`` (C ABI declared in ``include/drvish_b200.h``).

There is NO CPU fallback: if the shared library is missing or a CUDA device is absent the
product path raises (BASELINE.json north_star: "no CPU fallback").  Build it with
``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C drvish_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

DRV_MAX_LAYERS = 8
DRV_FLAG_FORCE_SIMT = 1
DRV_FLAG_NO_BN_UPDATE = 2
DRV_FLAG_STOP_AFTER_DECODER = 128
DRV_FLAG_ENCODER_BACKWARD_ONLY = 256
DRV_FLAG_LOSS_IN_GRADS = 2048
DRV_FLAG_FUSED_ALLREDUCE = 4096
DRV_STATUS_NAN_LOSS = 1
DRV_STATUS_MISSING_CLASS = 2
DRV_STATUS_BAD_LABEL = 4
DRV_GRAD_TAIL_FLOATS = 4  # loss + one float per status bit in the tail of the gradient buffer (data parallel)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrvish_b200.so")


class Dims(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int32),
        ("n_genes", C.c_int32),
        ("n_latent", C.c_int32),
        ("n_layers", C.c_int32),
        ("layers", C.c_int32 * DRV_MAX_LAYERS),
        ("n_drugs", C.c_int32),
        ("n_classes", C.c_int32),
        ("n_doses_total", C.c_int32),
    ]


class HParams(C.Structure):
    _fields_ = [
        ("scale_factor", C.c_float),
        ("epsilon", C.c_float),
        ("lib_loc", C.c_float),
        ("lib_scale", C.c_float),
        ("lam_scale", C.c_float),
        ("bias_scale", C.c_float),
        ("sigma_scale", C.c_float),
        ("alpha", C.c_float),
        ("bn_eps", C.c_float),
        ("bn_momentum", C.c_float),
        ("include_guide_factor", C.c_int32),
        ("flags", C.c_int32),
    ]


class StepOut(C.Structure):
    _fields_ = [("loss", C.c_double), ("status", C.c_uint32), ("path", C.c_uint32)]


# name -> (restype, argtypes); every symbol include/drvish_b200.h declares
_P, _I64P = C.c_void_p, C.POINTER(C.c_int64)
SYMBOLS = {
    "drv_version": (C.c_char_p, []),
    "drv_last_launch_count": (C.c_int, []),
    "drv_debug_w3_trace": (C.c_int, [C.POINTER(C.c_ulonglong), C.c_int]),
    "drv_param_offsets": (C.c_int, [C.POINTER(Dims), _I64P, _I64P, C.c_int, _I64P]),
    "drv_bn_offsets": (C.c_int, [C.POINTER(Dims), _I64P, _I64P, C.c_int, _I64P]),
    "drv_workspace_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "drv_elbo_step": (C.c_int, [C.POINTER(Dims), C.POINTER(HParams)] + [_P] * 15 + [C.c_size_t, _P]),
    "drv_encoder_forward": (C.c_int, [C.POINTER(Dims), C.POINTER(HParams)] + [_P] * 9 + [C.c_size_t, _P]),
    "drv_decoder_forward": (C.c_int, [C.POINTER(Dims), C.POINTER(HParams)] + [_P] * 8 + [C.c_size_t, _P]),
    "drv_logit_mean_sigmoid": (C.c_int, [C.POINTER(Dims)] + [_P] * 8 + [C.c_size_t, _P]),
    "drv_profile_enable": (C.c_int, [C.c_int]),
    "drv_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "drv_profile_read_kernels": (C.c_int, [C.POINTER(C.c_float)]),
    "drv_fill_normal": (C.c_int, [_P, C.c_int64, C.c_uint64, C.c_uint64, _P, C.c_uint64, _P]),
    "drv_fill_laplace": (C.c_int, [_P, C.c_int64, C.c_uint64, C.c_uint64, _P, C.c_uint64, _P]),
    "drv_counter_add": (C.c_int, [_P, C.c_uint64, _P]),
    "drv_host_pack_counts_u16": (C.c_int64, [_P, _P, C.c_int64, C.c_int]),
    "drv_unpack_counts_u16": (C.c_int, [_P, _P, C.c_int64, _P]),
    "drv_host_pack_counts_u8": (C.c_int64, [_P, _P, C.c_int64, _P, _P, C.c_int64, C.c_int]),
    "drv_host_pack_isa": (C.c_char_p, []),
    "drv_unpack_counts_u8": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int64, _P]),
    "drv_elbo_step_mcv": (C.c_int, [_P] * 14 + [C.c_size_t, _P]),
    "drv_csr_gather_rows": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "drv_split_molecules": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _P, _P, _P, _P, _P]),
    "drv_comm_header_bytes": (C.c_size_t, []),
    "drv_comm_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), _P]),
    "drv_comm_open": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "drv_comm_close": (C.c_int, [_P]),
    "drv_comm_free": (C.c_int, [_P]),
    "drv_peer_allreduce": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64, _P]),
    "drv_comm_bind": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int64]),
    "drv_shape_padded": (C.c_int, [C.POINTER(Dims)]),
    "drv_aggmo_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "drv_draw_step_noise": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P, C.c_uint64, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library; raises with build instructions when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"drvish_b200: {LIB_PATH} is missing - the CUDA extension is required (no CPU fallback). "
            "Build it with `make -C drvish_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code == 0:
        return
    if code == -1:
        raise RuntimeError(f"drvish_b200.{what}: invalid arguments (DRV_EINVAL)")
    if code == -2:
        raise RuntimeError(f"drvish_b200.{what}: workspace too small (DRV_ENOSPC)")
    raise RuntimeError(f"drvish_b200.{what}: CUDA error {code}")
