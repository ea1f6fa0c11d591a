import ctypes as C, torch, sys
sys.path.insert(0,'.')
from drvish_b200 import _lib
lib=_lib.load()
fn=lib.drv_test_tc_gemm; fn.restype=C.c_int
fn.argtypes=[C.c_void_p,C.c_int,C.c_int,C.c_void_p,C.c_int,C.c_int,C.c_void_p,C.c_int,C.c_int,C.c_int,C.c_int,C.c_int,C.c_void_p]
cfg=lib.drv_test_tc_gemm_cfg; cfg.argtypes=[C.c_int]*3
def run(a_mn,b_mn,M,N,K,splits=1):
    torch.manual_seed(0)
    Ad=torch.randn(M,K,device='cuda'); Bd=torch.randn(N,K,device='cuda')
    ref=(Ad.double()@Bd.double().t()).float()
    A=Ad.t().contiguous() if a_mn else Ad
    B=Bd.t().contiguous() if b_mn else Bd
    Cm=torch.zeros(M,N,device='cuda')
    code=fn(A.data_ptr(),A.stride(0),int(a_mn),B.data_ptr(),B.stride(0),int(b_mn),Cm.data_ptr(),N,M,N,K,splits,torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    err=float((Cm-ref).norm()/ref.norm())
    return code,err,float(Cm.abs().max()),Cm,ref
for (lbo,sbo,ks) in [(4096,512,1024),(512,4096,1024),(4096,1024,1024)]:
    cfg(lbo,sbo,ks)
    for (a,b) in [(True,False),(False,True),(True,True)]:
        for (M,N,K) in [(128,64,8),(128,64,32),(128,128,64)]:
            code,err,mx,Cm,ref=run(a,b,M,N,K)
            print(f"lbo={lbo} sbo={sbo} kstep={ks} a_mn={a} b_mn={b} M{M} N{N} K{K}: code={code} err={err:.4f} max={mx:.3f}")
