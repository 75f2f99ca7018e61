import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ocbnn-public_b200"))
from ocbnn_b200 import engine as E
g = torch.Generator().manual_seed(0)
for K in (32, 128, 512, 2048, 8192, 16384):
    for kind in ("randn", "positive"):
        A = torch.randn(256, K, generator=g); B = torch.randn(256, K, generator=g)
        if kind == "positive": A, B = A.abs(), B.abs()
        C = E.gemm_tf32x3(A.cuda(), B.cuda()).cpu().double()
        ref = A.double() @ B.double().T
        sc = A.double().abs() @ B.double().abs().T
        e = (C - ref) / sc
        f32 = (A.cuda() @ B.cuda().T).cpu().double()   # cuBLAS fp32 for comparison (may use TF32 if allowed; default off)
        e32 = (f32 - ref) / sc
        print(f"K={K:6d} {kind:8s} 3xTF32: max {e.abs().max():.2e} mean-signed {e.mean():+.2e} rms {e.pow(2).mean().sqrt():.2e} | fp32 cublas: max {e32.abs().max():.2e} rms {e32.pow(2).mean().sqrt():.2e}")
