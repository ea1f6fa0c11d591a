"""tcgen05/TMA GEMM building block against a plain PyTorch fp32 matmul (tf32 operand rounding:
10-bit mantissa -> 1e-3 relative per product, far less after accumulation)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _gemm(A, a_mn, B, b_mn, M, N, K, splits):
    from drvish_b200 import _lib

    lib = _lib.load()
    fn = lib.drv_test_tc_gemm
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                   C.c_int, C.c_int, C.c_void_p]
    Cm = torch.zeros(M, N, device="cuda")
    code = fn(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), Cm.data_ptr(), N, M, N, K,
              splits, torch.cuda.current_stream().cuda_stream)
    assert code == 0, code
    torch.cuda.synchronize()
    return Cm


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K,splits", [(128, 64, 32, 1), (256, 256, 256, 1), (384, 128, 1000, 3), (200, 96, 520, 2)])
def test_tc_gemm_matches_fp32(a_mn, b_mn, M, N, K, splits):
    torch.manual_seed(M + N + K)
    Ad = torch.randn(M, K, device="cuda")
    Bd = torch.randn(N, K, device="cuda")
    ref = (Ad.double() @ Bd.double().t()).float()
    A = Ad.t().contiguous() if a_mn else Ad
    B = Bd.t().contiguous() if b_mn else Bd
    out = _gemm(A, a_mn, B, b_mn, M, N, K, splits)
    err = (out - ref).norm() / ref.norm()
    assert err < 2e-3, float(err)
    # tf32 truncation reference: must agree much tighter than fp32-vs-tf32
    At = (Ad.view(torch.int32) & ~0x1FFF).view(torch.float32)
    Bt = (Bd.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref_t = (At.double() @ Bt.double().t()).float()
    err_t = (out - ref_t).norm() / ref_t.norm()
    assert err_t < 2e-5, float(err_t)


def _gemm_h(A, a_mn, B, b_mn, M, N, K, splits, alpha):
    from drvish_b200 import _lib

    lib = _lib.load()
    fn = lib.drv_test_tc_gemm_h
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                   C.c_int, C.c_int, C.c_float, C.c_void_p]
    Cm = torch.zeros(M, N, device="cuda")
    code = fn(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), Cm.data_ptr(), N, M, N, K,
              splits, alpha, torch.cuda.current_stream().cuda_stream)
    assert code == 0, code
    torch.cuda.synchronize()
    return Cm


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K,splits", [(128, 64, 64, 1), (256, 256, 256, 1), (384, 128, 1000, 3), (200, 192, 520, 2),
                                          (4096, 256, 10000, 0), (10000, 256, 4096, 0)])
def test_tc_gemm_fp16_operands(a_mn, b_mn, M, N, K, splits):
    """kind::f16 variant (the decoder's backward GEMMs): fp16 operands in both majors, fp32 accumulate,
    alpha rescaling; exact products of the fp16-rounded inputs up to accumulation order."""
    torch.manual_seed(M + N + K)
    Ad = torch.randn(M, K, device="cuda").half()
    Bd = torch.randn(N, K, device="cuda").half()
    ref = 0.25 * (Ad.double() @ Bd.double().t()).float()
    A = Ad.t().contiguous() if a_mn else Ad
    B = Bd.t().contiguous() if b_mn else Bd
    out = _gemm_h(A, a_mn, B, b_mn, M, N, K, splits, 0.25)
    err = (out - ref).norm() / ref.norm()
    assert err < 2e-5, float(err)
