"""GPU: the TMA + tcgen05 contraction kernel against fp64 torch on the same seeded inputs.
backend 2 = the product path, error-compensated 3xTF32: must match fp64 to fp32 round-off (2e-6 * sqrt(K) * scale).
backend 3 = single-pass kind::tf32 (measurement only): the tensor core truncates both operands to 10 mantissa bits,
products accumulate in fp32; raw fp32 inputs are allowed 6e-3 * sqrt(K) * rms(a) * rms(b).  Layout / descriptor / pipeline correctness is pinned EXACTLY: operands
quantised to multiples of 1/64 in [-4, 4] are TF32-representable, every product is a multiple of 2^-12 and every
partial sum stays below 2^12, so fp32 accumulation is exact in any order and the result must equal fp64 torch."""
import importlib
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import PKG_NAME

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    return importlib.import_module(PKG_NAME + '.ops')


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=g) * scale


def quant(t):
    return ((t * 64).round() / 64).clamp(-4, 4)


def tf32_close(C, ref, K, what, ascale=1.0, bscale=1.0):
    C, ref = C.detach().double().cpu(), ref.double().cpu()
    assert C.shape == ref.shape
    err = (C - ref).abs().max().item()
    tol = 6e-3 * math.sqrt(max(K, 1)) * ascale * bscale + 1e-6
    assert err <= tol, f'{what}: max err {err:.3e} > {tol:.3e}'


# shapes of the LeWin / LeFF / head path: ragged K (28, 56), ragged M and N, deep K, all four operand layouts
SHAPES = [(128, 32, 32), (256, 56, 56), (4096, 224, 56), (1000, 112, 28), (2048, 448, 3584), (300, 96, 448),
          (64, 16, 8), (129, 33 * 4, 200), (1024, 65536 // 16, 448), (512, 8, 72), (512, 27, 144), (516, 24, 72),
          (40, 12, 520), (8, 8, 8)]


@pytest.mark.parametrize('backend', [2, 3, 4, 5])
@pytest.mark.parametrize('M,N,K', SHAPES)
@pytest.mark.parametrize('tA,tB', [(False, True), (False, False), (True, False), (True, True)])
def test_gemm_tc_plain(ops, M, N, K, tA, tB, backend):
    if (tA and M % 4) or ((not tA) and K % 4) or (tB and K % 4) or ((not tB) and N % 4):    # lda / ldb pitch
        pytest.skip('row pitch not 16-byte aligned: SIMT path covers it')
    A = quant(gen(K, M) if tA else gen(M, K))
    B = quant(gen(N, K, seed=1) if tB else gen(K, N, seed=1))
    C = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm(A.cuda(), B.cuda(), C, transA=tA, transB=tB, backend=backend)
    ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double()
    assert ref.abs().max().item() < 4096
    err = (C.double().cpu() - ref).abs().max().item()
    assert err == 0.0, f'tc gemm {M}x{N}x{K} tA={tA} tB={tB}: max err {err:.3e} on exactly representable data'


@pytest.mark.parametrize('M,N,K,tA,tB', [(4096, 224, 56, False, True), (2048, 448, 3584, False, False),
                                         (1024, 4096, 448, False, True), (448, 224, 8192, True, False)])
def test_gemm_tc_fp32_inputs(ops, M, N, K, tA, tB):
    A = gen(K, M) if tA else gen(M, K)
    B = gen(N, K, seed=1) if tB else gen(K, N, seed=1)
    ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double()
    C = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm(A.cuda(), B.cuda(), C, transA=tA, transB=tB, backend=3)
    tf32_close(C, ref, K, f'tc gemm 1xTF32 {M}x{N}x{K} tA={tA} tB={tB}')
    # product path: fp32-level accuracy, and no worse than the exact-fp32 SIMT kernel by more than 2x
    C3 = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm(A.cuda(), B.cuda(), C3, transA=tA, transB=tB, backend=2)
    C1 = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm(A.cuda(), B.cuda(), C1, transA=tA, transB=tB, backend=1)
    e3 = (C3.double().cpu() - ref).abs().max().item()
    e1 = (C1.double().cpu() - ref).abs().max().item()
    assert e3 <= 3 * e1 + 1e-6, f'3xTF32 {M}x{N}x{K}: max err {e3:.3e} (fp32 SIMT {e1:.3e})'
    print(f'3xTF32 err {e3:.3e}, fp32 SIMT err {e1:.3e}')
    # reduced-pass modes against their DEFINITION (same operand treatment in fp64): 2x = A exact, B rounded to nearest
    # TF32; 1x = both rounded to nearest.  What is left is accumulation round-off, i.e. the 3x-level bound.
    def rn(x):
        return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)
    opA, opB = (A.t() if tA else A), (B.t() if tB else B)
    for backend, ra, rb in ((4, opA, rn(opB)), (5, rn(opA), rn(opB))):
        Cx = torch.full((M, N), float('nan'), device='cuda')
        ops.gemm(A.cuda(), B.cuda(), Cx, transA=tA, transB=tB, backend=backend)
        refx = ra.double() @ rb.double()
        ex = (Cx.double().cpu() - refx).abs().max().item()
        assert ex <= 3 * e1 + 1e-6, f'backend {backend} {M}x{N}x{K}: {ex:.3e} away from its own definition (fp32 SIMT {e1:.3e})'
        # and the operand rounding itself stays inside the TF32 bound (unbiased: ~2^-12 per operand, sqrt(K) growth)
        tf32_close(Cx, ref, K, f'backend {backend} vs fp64', ascale=0.5)
        # pre-rounded B + b_is_tf32: the kernel skips B and must give the same result bit for bit
        Cy = torch.full((M, N), float('nan'), device='cuda')
        ops.gemm(A.cuda(), rn(B).cuda(), Cy, transA=tA, transB=tB, backend=backend, b_is_tf32=True)
        # (bit-identical unless the reduction is split over CTAs and summed with atomics in arrival order)
        dxy = (Cx - Cy).abs().max().item()
        assert dxy == 0.0 or (K >= 512 and dxy <= 4e-6 * ref.abs().max().item()), f'backend {backend}: b_is_tf32 path differs from in-kernel rounding by {dxy:.3e}'


def test_gemm_tc_epilogues(ops):
    M, N, K = 320, 224, 56
    A, W, b = gen(M, K), gen(N, K, seed=1), gen(N, seed=2)
    R, aux = gen(M, N, seed=3), gen(M, N, seed=4)
    rs = torch.tensor([0.0, 1.0 / 0.9, 1.0 / 0.9, 1.0, 0.5])
    base = (A.double() @ W.double().t() + b.double()).float()
    dA, dW, db = A.cuda(), W.cuda(), b.cuda()
    C = torch.empty(M, N, device='cuda'); pre = torch.empty(M, N, device='cuda')
    ops.gemm(dA, dW, C, bias=db, act=ops.ACT_GELU, preact=pre, backend=2)
    tf32_close(pre, base, K, 'preact'); tf32_close(C, F.gelu(base), K, 'gelu')
    ops.gemm(dA, dW, C, bias=db, act=ops.ACT_LRELU, act_param=0.1, backend=2)
    tf32_close(C, F.leaky_relu(base, 0.1), K, 'lrelu')
    ops.gemm(dA, dW, C, bias=db, rowscale=rs.cuda(), rows_per_scale=64, residual=R.cuda(), backend=2)
    tf32_close(C, R + rs.repeat_interleave(64)[:, None] * base, K, 'residual+rowscale')
    ag = aux.clone().requires_grad_(True)
    F.gelu(ag).sum().backward()
    ops.gemm(dA, dW, C, aux=aux.cuda(), aux_act=ops.ACT_GELU, backend=2)
    tf32_close(C, (A @ W.t()) * ag.grad, K, 'dgelu')
    ops.gemm(dA, dW, C, aux=aux.cuda(), aux_act=ops.ACT_MUL, backend=2)        # aux already holds the derivative
    tf32_close(C, (A @ W.t()) * aux, K, 'aux multiply')
    big = torch.zeros(M, 2 * N, device='cuda'); big[:, N:] = R.cuda()
    ops.gemm(dA, dW, big[:, N:], accumulate=True, alpha=0.5, backend=2)
    tf32_close(big[:, N:], R + 0.5 * (A @ W.t()), K, 'accumulate strided')
    assert big[:, :N].abs().max().item() == 0


def test_gemm_tc_splitk_weight_grad(ops):
    M, N, K = 40000, 56, 224          # dW[N,K] = dY^T[N,M] X[M,K], reduction over M (split-K + fp32 atomics)
    dY, X = gen(M, N, scale=0.1), gen(M, K, seed=1)
    G0 = gen(N, K, seed=2)
    G = G0.cuda()
    ops.gemm(dY.cuda(), X.cuda(), G, transA=True, transB=False, accumulate=True, backend=2)
    tf32_close(G, G0 + (dY.double().t() @ X.double()).float(), M, 'split-k', ascale=0.1)


def test_gemm_tc_matches_simt_on_tf32_exact_inputs(ops):
    """Inputs already representable in TF32 (10-bit mantissa): both kernels must agree to fp32 round-off."""
    M, N, K = 512, 128, 256
    A, B = quant(gen(M, K)), quant(gen(N, K, seed=1))
    C1 = torch.empty(M, N, device='cuda'); C2 = torch.empty(M, N, device='cuda')
    ops.gemm(A.cuda(), B.cuda(), C1, backend=1)
    ops.gemm(A.cuda(), B.cuda(), C2, backend=2)
    assert (C1 - C2).abs().max().item() == 0.0


@pytest.mark.parametrize('backend', [0, 1, 2])
@pytest.mark.parametrize('M,N,K,tA', [(224, 56, 40000, True), (448, 448, 4096, True), (56, 224, 3000, True), (300, 64, 96, False)])
def test_gemm_a_rowsum(ops, M, N, K, tA, backend):
    """FaGemmEpilogue.a_rowsum: the bias gradient (column sums of the stored dY when transA) from the same pass."""
    A = gen(K, M, scale=0.1) if tA else gen(M, K, scale=0.1)
    B = gen(K, N, seed=1)
    C0 = gen(M, N, seed=2)
    rs0 = gen(M, seed=3)
    C, rs = C0.cuda(), rs0.cuda()
    ops.gemm(A.cuda(), B.cuda(), C, transA=tA, transB=False, accumulate=True, a_rowsum=rs, backend=backend)
    opA = (A.t() if tA else A).double()
    ref = C0.double() + opA @ B.double()
    tol = 3e-6 * math.sqrt(K) * 0.1 * 4 + 1e-5
    assert (C.double().cpu() - ref).abs().max().item() <= tol * 10
    ref_rs = rs0.double() + opA.sum(1)
    assert (rs.double().cpu() - ref_rs).abs().max().item() <= 2e-6 * K ** 0.5 * 0.1 * 8 + 1e-5


@pytest.mark.parametrize('backend', [0, 2])
def test_gemm_a_kscale_and_bwd_rowscale(ops, backend):
    """DropPath backward folded into the contractions: dW += (D dY)^T X with D = per-sample scale along the reduction
    (FaGemmEpilogue.a_kscale, also applied to a_rowsum), and dX = D (dY W) * gelu'(aux) with the row scale in the BWD
    epilogue."""
    T, Co, Ci, rps = 6 * 64, 56, 224, 64
    dY, X = gen(T, Co, scale=0.1), gen(T, Ci, seed=1)
    dp = torch.tensor([0.0, 1 / 0.9, 1 / 0.9, 0.0, 1 / 0.9, 1 / 0.9])
    sc = dp.repeat_interleave(rps)[:, None]
    G0, b0 = gen(Co, Ci, seed=2), gen(Co, seed=3)
    G, bsum = G0.cuda(), b0.cuda()
    ops.gemm(dY.cuda(), X.cuda(), G, transA=True, transB=False, accumulate=True, a_rowsum=bsum, a_kscale=dp.cuda(),
             a_k_rows_per_scale=rps, backend=backend)
    ref = G0.double() + (dY * sc).double().t() @ X.double()
    assert (G.double().cpu() - ref).abs().max().item() <= 2e-5
    assert (bsum.double().cpu() - (b0.double() + (dY * sc).double().sum(0))).abs().max().item() <= 2e-5
    W, aux = gen(Co, Ci, seed=4, scale=0.1), gen(T, Ci, seed=5)
    dX = torch.empty(T, Ci, device='cuda')
    ops.gemm(dY.cuda(), W.cuda(), dX, transB=False, aux=aux.cuda(), aux_act=ops.ACT_GELU, rowscale=dp.cuda(),
             rows_per_scale=rps, backend=backend)
    ag = aux.clone().double().requires_grad_(True)
    F.gelu(ag).sum().backward()
    refx = (dY * sc).double() @ W.double() * ag.grad
    assert (dX.double().cpu() - refx).abs().max().item() <= 2e-5
    dX2 = torch.empty(T, Ci, device='cuda')
    ops.gemm(dY.cuda(), W.cuda(), dX2, transB=False, rowscale=dp.cuda(), rows_per_scale=rps, backend=backend)
    assert (dX2.double().cpu() - (dY * sc).double() @ W.double()).abs().max().item() <= 2e-5
    with pytest.raises(RuntimeError):            # no silent fallback: the SIMT kernel has no a_kscale
        ops.gemm(dY.cuda(), X.cuda(), G, transA=True, transB=False, accumulate=True, a_kscale=dp.cuda(),
                 a_k_rows_per_scale=rps, backend=1)


@pytest.mark.parametrize('backend', [2, 3, 5])
@pytest.mark.parametrize('tA,tB', [(False, True), (False, False), (True, False), (True, True)])
@pytest.mark.parametrize('M,N,K', [(128 * 151 + 40, 256, 96), (128 * 150, 64, 1056), (128 * 3 + 4, 128 * 60, 72),
                                   (128 * 1200 + 8, 28, 28), (128 * 700, 56, 56), (128 * 500 + 4, 96, 28)])
def test_gemm_tc_many_tiles(ops, M, N, K, tA, tB, backend):
    """Shapes with several waves of tiles per SM (persistent loop, ring and accumulator phases wrapping many times), odd
    tile counts, ragged M, both B layouts, split-K; exact on TF32-representable data.  The last three are the one-k-block
    tiles of the 128 x 128 level, where the A ring in TMEM is 7 / 6 / 5 stages deep (32- / 64- / 96-wide tiles)."""
    A = quant(gen(K, M) if tA else gen(M, K))
    B = quant(gen(N, K, seed=1) if tB else gen(K, N, seed=1))
    C = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm(A.cuda(), B.cuda(), C, transA=tA, transB=tB, backend=backend)
    ref = (A.t() if tA else A).double() @ (B.t() if tB else B).double()
    err = (C.double().cpu() - ref).abs().max().item()
    assert err == 0.0, f'tc gemm {M}x{N}x{K} tA={tA} tB={tB}: max err {err:.3e}'
