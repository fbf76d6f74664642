"""One window-attention configuration launched a few times (ncu target: -k regex:win_attn -s 2 -c 2).
Usage: python tools/one_attn.py [H=64] [heads=4]"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = 'frequency-wised_all-in-one_image_restoration_model_b200'
ops = importlib.import_module(PKG + '.ops')
fd = importlib.import_module(PKG + '.net.utils.frequency_decompose')
H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B, hd, shift = 16, 56, 4
C = heads * hd; T = B * H * H
qkv = torch.randn(T, 3 * C, device='cuda') * 0.5
table = torch.randn(225, heads, device='cuda') * 0.3
coef = torch.randn(B, heads, 3, device='cuda') * 0.3
bob = fd.half_band_map('frequency_decompose_1', 0.5, 64).cuda()
o = torch.empty(T, C, device='cuda'); dO = torch.randn(T, C, device='cuda')
dq = torch.empty(T, C, device='cuda'); dkv = torch.empty(T, 2 * C, device='cuda')
dtab = torch.zeros_like(table); dcf = torch.zeros_like(coef)
args = (B, H, H, heads, hd, shift, hd ** -0.5)
for _ in range(3):
    ops.win_attn_fwd(qkv[:, :C], qkv[:, C:], o, *args, table, coef, heads, bob, 3)
    ops.win_attn_bwd(qkv[:, :C], qkv[:, C:], dO, dq, dkv, *args, table, dtab, coef, heads, dcf, bob, 3)
torch.cuda.synchronize()
print('ok', float(o.abs().mean()))
