"""Microbenchmark: the decoder's two backward GEMM shapes with tf32 (fp32 storage) vs fp16 operands.
Run on the GPU box: python tools/gemm_h_bench.py [B G h]"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from drvish_b200 import _lib  # noqa: E402

B, G, h = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 5000, 256)
lib = _lib.load()
f32 = lib.drv_test_tc_gemm
f32.restype = C.c_int
f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
f16 = lib.drv_test_tc_gemm_h
f16.restype = C.c_int
f16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]
st = torch.cuda.current_stream().cuda_stream
NCOPY = 4
for dt, name in ((torch.float32, "tf32"), (torch.float16, "fp16")):
    dl = [torch.randn(B, 2 * G, device="cuda").to(dt) for _ in range(NCOPY)]
    W = torch.randn(2 * G, h, device="cuda").to(dt)
    act = torch.randn(B, h, device="cuda").to(dt)
    dA = torch.zeros(B, h, device="cuda")
    dW = torch.zeros(2 * G, h, device="cuda")

    def run_dact(i):
        d = dl[i % NCOPY]
        if dt == torch.float32:
            return f32(d.data_ptr(), 2 * G, 0, W.data_ptr(), h, 1, dA.data_ptr(), h, B, h, 2 * G, 0, st)
        return f16(d.data_ptr(), 2 * G, 0, W.data_ptr(), h, 1, dA.data_ptr(), h, B, h, 2 * G, 0, 1.0, st)

    def run_dw(i):
        d = dl[i % NCOPY]
        if dt == torch.float32:
            return f32(d.data_ptr(), 2 * G, 1, act.data_ptr(), h, 1, dW.data_ptr(), h, 2 * G, h, B, 0, st)
        return f16(d.data_ptr(), 2 * G, 1, act.data_ptr(), h, 1, dW.data_ptr(), h, 2 * G, h, B, 0, 1.0, st)

    for fn, what in ((run_dact, "dAct"), (run_dw, "dW")):
        for i in range(3):
            assert fn(i) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{name} {what}: {ms * 1e3:.1f} us  ({2.0 * B * 2 * G * h / ms / 1e9:.0f} TFLOP/s) [includes the memset of a split-K target]")
