"""One fa_gemm configuration launched a few times (ncu target: -k regex:gemm_tc -s 2 -c 1).
Usage: python tools/one_gemm.py M N K tA tB backend [bexact] [epi]   epi in {plain, gelu (bias+GELU+preact), res (bias+residual)}"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = 'frequency-wised_all-in-one_image_restoration_model_b200'
ops = importlib.import_module(PKG + '.ops')
M, N, K, tA, tB, backend = (int(v) for v in sys.argv[1:7])
bexact = len(sys.argv) > 7 and sys.argv[7] == '1'
epi = sys.argv[8] if len(sys.argv) > 8 else 'plain'
A = torch.randn((K, M) if tA else (M, K), device='cuda')
B = torch.randn((N, K) if tB else (K, N), device='cuda') * 0.05
C = torch.zeros(M, N, device='cuda')
kw = {}
if epi == 'gelu':
    kw = dict(bias=torch.randn(N, device='cuda'), act=ops.ACT_GELU, preact=torch.empty(M, N, device='cuda'))
elif epi == 'res':
    kw = dict(bias=torch.randn(N, device='cuda'), residual=torch.randn(M, N, device='cuda'))
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(4):
    flush.zero_()
    ops.gemm(A, B, C, transA=bool(tA), transB=bool(tB), backend=backend, b_is_tf32=bexact, **kw)
torch.cuda.synchronize()
print('ok', float(C.abs().mean()))
