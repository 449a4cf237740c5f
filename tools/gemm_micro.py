import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from micro_diffusion_b200.ops import CudaOps
o = CudaOps("cuda:0"); dev = torch.device("cuda:0")
def bench(M,N,K,epi=0,iters=30):
    A=torch.randn(M,K,device=dev).bfloat16(); B=torch.randn(N,K,device=dev).bfloat16()
    Cm=torch.zeros(M,N,device=dev,dtype=torch.bfloat16 if epi==0 else torch.float32)
    res=torch.zeros(M,N,device=dev) if epi==2 else None
    for _ in range(3): o.gemm(A,B,Cm,epi=epi,res=res)
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): o.gemm(A,B,Cm,epi=epi,res=res)
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/iters
    print(f"M={M} N={N} K={K} epi={epi}: {ms*1e3:.1f} us {2*M*N*K/ms/1e9:.0f} TF/s", flush=True)
print("MD_GEMM_DEBUG=", os.environ.get("MD_GEMM_DEBUG"), "MD_GEMM_EPI=", os.environ.get("MD_GEMM_EPI"))
for K in (256, 1024, 4096):
    bench(16384, 1024, K); bench(16384, 3072, K)
bench(16384,1024,1024,epi=2)
