"""CPU checks of the C-ABI boundary: the shared library loads, exports every symbol that
include/drvish_b200.h declares, and its host-only entry points (layouts, workspace sizing)
agree with the Python mirror.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

from drvish_b200 import _lib
from drvish_b200.runtime import make_dims

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "drvish_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(drv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/drvish_b200.h but not exported"
    # and the ctypes mirror binds every declared symbol
    assert set(names) <= set(_lib.SYMBOLS), set(names) - set(_lib.SYMBOLS)
    assert lib.drv_version().decode().startswith("drvish_b200")


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Dims) == 4 * (4 + _lib.DRV_MAX_LAYERS + 3)
    assert C.sizeof(_lib.HParams) == 4 * 12
    assert C.sizeof(_lib.StepOut) == 16


def test_param_layout_matches_state_dict_order():
    import drvish_b200 as D

    lib = _lib.load()
    model = D.DRNBVAE(n_input=37, n_classes=3, n_drugs=2, n_conditions=[4, 6], n_latent=5, layers=[16, 12, 8])
    d = make_dims(16, 37, 5, [16, 12, 8], 2, 3, 10)
    n = 64
    offs, sizes, total = (C.c_int64 * n)(), (C.c_int64 * n)(), C.c_int64()
    cnt = lib.drv_param_offsets(C.byref(d), offs, sizes, n, C.byref(total))
    named = [(k, v) for k, v in model.named_parameters()]
    enc_dec = [(k, v) for k, v in named if not k.startswith("lmb_list")]
    assert cnt == len(enc_dec) + 4
    prev_end = 0
    for i, (k, v) in enumerate(enc_dec):
        assert sizes[i] == v.numel(), k
        assert offs[i] >= prev_end and offs[i] % 32 == 0, k  # 128-byte aligned, non-overlapping
        prev_end = offs[i] + sizes[i]
    assert sizes[len(enc_dec)] == 2 * 5 and sizes[len(enc_dec) + 2] == 10
    assert total.value >= prev_end
    nb = 64
    boffs, bsizes, btotal = (C.c_int64 * nb)(), (C.c_int64 * nb)(), C.c_int64()
    bcnt = lib.drv_bn_offsets(C.byref(d), boffs, bsizes, nb, C.byref(btotal))
    n_bn = sum(isinstance(m, torch.nn.BatchNorm1d) for m in model.modules())
    assert bcnt == 2 * n_bn
    assert lib.drv_workspace_bytes(C.byref(d)) > 0


def test_invalid_dims_are_rejected():
    lib = _lib.load()
    d = make_dims(1, 37, 5, [16], 0, 0, 0)  # a single cell: BatchNorm needs >= 2
    assert lib.drv_workspace_bytes(C.byref(d)) == 0
    assert lib.drv_param_offsets(C.byref(d), None, None, 0, None) == -1
    with pytest.raises(ValueError):
        make_dims(8, 37, 5, [], 0, 0, 0)
