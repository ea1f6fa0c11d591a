import ctypes as C, sys, torch
sys.path.insert(0, '.')
from drvish_b200 import _lib
lib = _lib.load()
fn = lib.drv_test_tc_gemm; fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
def rna(t):
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
def gemm(A, a_mn, B, b_mn, M, N, K):
    Cm = torch.zeros(M, N, device='cuda')
    code = fn(A.data_ptr(), A.stride(0), int(a_mn), B.data_ptr(), B.stride(0), int(b_mn), Cm.data_ptr(), N, M, N, K, 0, torch.cuda.current_stream().cuda_stream)
    assert code == 0
    torch.cuda.synchronize()
    return Cm
torch.manual_seed(0)
Bc, G2, h = 4096, 10000, 256
# realistic-ish operands: dlogits with heavy tails and a column mean, W small
d = (torch.randn(Bc, G2, device='cuda') * torch.rand(1, G2, device='cuda') + 0.3 * torch.randn(1, G2, device='cuda'))
W = torch.randn(G2, h, device='cuda') * 0.06
act = torch.relu(torch.randn(Bc, h, device='cuda'))
dr, Wr, ar = rna(d), rna(W), rna(act)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
# dAct = d @ W   (A K-major, B MN-major)
ex_r = (dr.double() @ Wr.double())
ex = (d.double() @ W.double())
out = gemm(dr, False, Wr, True, Bc, h, G2)
print('dAct: gemm vs exact(rounded ops) %.2e | rounded vs unrounded %.2e | column-mean err of gemm %.2e' % (rel(out, ex_r), rel(ex_r, ex), rel(out.mean(0), ex_r.mean(0))))
# dW = d^T @ act (A MN-major, B MN-major)
ex_r = (dr.double().t() @ ar.double()); ex = d.double().t() @ act.double()
out = gemm(dr, True, ar, True, G2, h, Bc)
print('dW:   gemm vs exact(rounded ops) %.2e | rounded vs unrounded %.2e' % (rel(out, ex_r), rel(ex_r, ex)))
# forward logits = act @ W^T with W as [2G][h] K-major
ex_r = ar.double() @ Wr.double().t(); ex = act.double() @ W.double().t()
out = gemm(ar, False, Wr, False, Bc, G2, h)
print('fwd:  gemm vs exact(rounded ops) %.2e | rounded vs unrounded %.2e' % (rel(out, ex_r), rel(ex_r, ex)))
