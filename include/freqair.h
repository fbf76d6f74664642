/*
 * freqair.h - C ABI of libfreqair.so, the B200 (sm_100a) kernels behind the restoration-network
 * forward/backward hot path of stcodeer/Frequency-wised_All-in-One_Image_Restoration_Model.
 *
 * The reference has no FFI of its own (it is 100 % Python; every GPU op is a torch library call), so
 * each entry point below names the reference call site(s) it replaces ("ref:" = path:line under the
 * reference tree).  Conventions:
 *   - plain pointers + sizes, no torch types; all tensors are fp32, contiguous unless an ld* is given;
 *   - the caller owns every buffer (inputs, outputs, saved-for-backward, workspaces); the library
 *     never allocates, frees or retains device memory;
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*) and returns;
 *   - return 0 on success, non-zero on error with the text in fa_last_error_string();
 *     unsupported shapes are errors, never silent fallbacks.  There is no CPU path.
 * "tokens" layout = [B, H*W, C] row-major (NHWC), the layout net/*_Uformer.py keeps between layers.
 */
#ifndef FREQAIR_H
#define FREQAIR_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fa_stream_t; /* cudaStream_t */

/* ------------------------------------------------------------------ library */
#define FREQAIR_ABI_VERSION 6      /* bumped whenever a prototype or struct in this header changes */
const char* fa_version(void);
int fa_abi_version(void);          /* the FREQAIR_ABI_VERSION the library was compiled against (checked at load time) */
const char* fa_last_error_string(void);
int fa_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* in-stream timing of one kernel class (bench.py roofline leg): cls = FA_K_* id, 0 = off */
int fa_prof_begin(int kernel_class);
int fa_prof_end(double* total_ms, int64_t* launches);
int64_t fa_launch_count(void);      /* kernels launched by this library since load / last reset */
void fa_launch_count_reset(void);

/* activation ids used by epilogues */
enum { FA_ACT_NONE = 0, FA_ACT_GELU = 1, FA_ACT_LRELU = 2, FA_ACT_SIGMOID = 3 };

/* ------------------------------------------------------------------ dense contraction (K3)
 * C[M,N] = epi(alpha * op(A)[M,K] * op(B)[K,N]);  op(A) = transA ? A[k*lda+m] : A[m*lda+k];
 * op(B) = transB ? B[n*ldb+k] : B[k*ldb+n]  (transB=1 is the nn.Linear weight layout [N,K]).
 * epi(v): v += bias[n]; (preact[m,n] = v); v = act(v); v *= act'(aux[m,n]); v *= rowscale[m / rows_per_scale];
 *         v += residual[m,n]; v += C[m,n] if accumulate.
 * ref: every nn.Linear on the path - decoder_Uformer.py:121-122,294; encoder_Uformer.py:96-97,305;
 *      leff.py:98,114; encoder_Uformer.py:975 (448->65536 head); encoder_ViT.py:77,97; and the
 *      autograd backward of each (cuBLAS sgemm in the reference).
 * backend: 0 = auto (tcgen05 error-compensated 3xTF32 - fp32-level accuracy - when the shape is eligible, else fp32
 *          SIMT), 1 = force fp32 SIMT, 2 = force tcgen05 3xTF32 (error if not eligible), 3 = force tcgen05 single-pass
 *          TF32 on the raw operands (truncated to 10 mantissa bits by the tensor core; measurement / comparison only),
 *          4 = tcgen05 2xTF32: op(A) exact (hi + lo split), op(B) rounded to the nearest TF32 inside the kernel
 *          (relative operand error <= 2^-12, unbiased), 5 = tcgen05 1xTF32 with BOTH operands rounded to nearest.
 *          4 and 5 fall back to the SIMT kernel for ineligible shapes like 0 does.  Which layer classes may use them is
 *          a parity decision made by the caller (the restorer's LeFF forward and the backward
 *          contractions of the Uformer path run 5; every other forward contraction stays 3x). */
typedef struct FaGemmEpilogue {
  const float* bias;
  int act; float act_param;
  const float* aux; int64_t ldaux; int aux_act; float aux_param;
  const float* rowscale; int rows_per_scale;
  const float* residual; int64_t ldr;
  int accumulate;
  float alpha;
  float* preact; int64_t ldpre;   /* if set: preact[m,n] = alpha*acc + bias[n] (value before act), for the backward */
  float* a_rowsum;                /* if set: a_rowsum[m] += sum_k op(A)[m,k].  With transA = 1 (dW = dY^T X) this is the
                                     column sum of the stored dY, i.e. the bias gradient, taken while the A tiles pass
                                     through the kernel instead of in a second pass over dY (fa_colsum) */
  /* if a_kscale is set: op(A)[m,k] is multiplied by a_kscale[k / a_k_rows_per_scale] on its way into the contraction
     (and into a_rowsum).  With transA = 1 and k = token index this is the per-sample DropPath scale of timm's DropPath
     backward folded into dW = (D dY)^T X, so the scaled gradient is never materialised.  a_k_rows_per_scale % 32 == 0;
     tcgen05 path only (an ineligible shape is an error, not a fallback). */
  const float* a_kscale;
  int a_k_rows_per_scale;
  int b_is_tf32;                  /* backends 4 / 5: op(B) is already TF32-representable (e.g. weights rounded once per
                                     optimiser step by fa_round_tf32): the kernel skips its in-place rounding of B */
} FaGemmEpilogue;
int fa_gemm(const float* A, const float* B, float* C, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc,
            int transA, int transB, const FaGemmEpilogue* epi, int backend, fa_stream_t stream);
/* out[n] (+)= sum_m rowscale[m/rps] * X[m*ld+n]   (bias gradients) */
int fa_colsum(const float* X, float* out, int M, int N, int64_t ld, const float* rowscale, int rows_per_scale,
              int accumulate, fa_stream_t stream);

/* ------------------------------------------------------------------ frequency-band decomposition (K1)
 * band_of_bin: uint8 [n][n/2+1], band id of each un-shifted rfft2 bin (built on the host with the
 * reference's own formula, frequency_decompose.py:17-26,38-48,80-88).
 * fa_band_split: y[band][map] = irfft2(rfft2(x[map]) * [band_of_bin == band])             (mode 0)
 *                y[band][map][n][n][2] = full un-shifted spectrum * mask, (re,im)          (mode 1)
 *   ref: FrequencyDecompose.frequency_decompose / _1, inverse=True / False (frequency_decompose.py:28-107)
 * fa_band_filter: y[map] = irfft2(rfft2(x[map]) * (1 + coef[g(map)][band_of_bin])) - the collapsed form of
 *   attn + sum_i lambda_i * band_i(attn) (decoder_Uformer.py:275-288, encoder_ViT.py:85-92); self-adjoint,
 *   so the same call is its own backward w.r.t. x.  g(map) = (map / maps_per_group) * heads + map % heads
 *   with coef [groups*heads][nbands].
 * fa_band_energy: out[g][band] += sum over maps of <a, band(b)> (gradient of the filter w.r.t. coef).
 * n in {8,16,32,64,128}. */
int fa_band_split(const float* x, float* y, int64_t nmaps, int n, const uint8_t* band_of_bin, int nbands, int mode,
                  fa_stream_t stream);
int fa_band_filter(const float* x, float* y, int64_t nmaps, int n, const uint8_t* band_of_bin, int nbands,
                   const float* coef, int maps_per_group, int heads, fa_stream_t stream);
int fa_band_energy(const float* a, const float* b, float* out, int64_t nmaps, int n, const uint8_t* band_of_bin,
                   int nbands, int maps_per_group, int heads, fa_stream_t stream);
/* Spectral L1 term of the loss: loss[0] += mean| D(a) - D(b) | with D = FrequencyDecompose('frequency_decompose',
 * 1/nbands, n, n, inverse=False), i.e. the mean over the [nbands, maps, n, n, 2] (re, im) stack; grad (may be NULL) =
 * gscale * dloss/da.  The bands partition the spectrum, so the stack is never materialised: one fft2 of a - b per map.
 * ref: train.py:69-70,90-91 + frequency_decompose.py:28-68. */
int fa_spectral_l1(const float* a, const float* b, float* loss, float* grad, int64_t nmaps, int n,
                   const uint8_t* band_of_bin, int nbands, float gscale, fa_stream_t stream);
/* mean / residual split (frequency_decompose.py:109-118): y[0]=mean broadcast, y[1]=x-mean */
int fa_dc_split(const float* x, float* y, int64_t nmaps, int n, fa_stream_t stream);

/* ------------------------------------------------------------------ LeWin window attention (K2)
 * q [B*H*W, ldq] and kv [B*H*W, ldkv] (k at column h*hd+d, v at C + h*hd+d) are in image token order; the
 * kernel gathers each 8x8 window (cyclic shift `shift`, decoder_Uformer.py:675-685) itself and scatters o
 * [B*H*W, C] back, so roll / window_partition / window_reverse never touch HBM.
 * S = scale*q.k^T + table[rel_index] (+ 0/-100 shift mask, :634-651); P = softmax(S);
 * P' = P filtered by (1 + coef[b][head][band]) if coef != NULL (:275-288); o = P'.v.
 * hd in {28, 56, 64}; table [225][heads] or NULL; coef [B][heads][nbands] or NULL.
 * ref: WindowAttention.forward decoder_Uformer.py:235-299, encoder_Uformer.py:152-183 ('origin' MSA),
 *      Attention.forward encoder_ViT.py:76-98 (H=W=8, table=NULL, shift=0, coef [B or 1][heads][nb]).
 * bwd recomputes P; dtable / dcoef are ACCUMULATED (caller zero-fills). coef_bstride = 0 shares coef over b.
 * drop_seed != NULL (a DEVICE pointer to one int64): nn.Dropout(drop_p) on P' before P'.v (encoder_ViT.py:94, train
 * mode): element (item, i, j) is kept with probability 1 - drop_p and scaled by 1 / (1 - drop_p), the mask being a
 * stateless hash of (*drop_seed, item, i, j) that the backward regenerates from the same seed.  Needs coef != NULL. */
int fa_win_attn_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, float* o, int B, int H, int W, int heads,
                    int hd, int shift, float scale, const float* table, const float* coef, int coef_bstride,
                    const uint8_t* band_of_bin, int nbands, float drop_p, const int64_t* drop_seed, fa_stream_t stream);
int fa_win_attn_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* dout, float* dq, float* dkv,
                    int B, int H, int W, int heads, int hd, int shift, float scale, const float* table, float* dtable,
                    const float* coef, int coef_bstride, float* dcoef, const uint8_t* band_of_bin, int nbands,
                    float drop_p, const int64_t* drop_seed, fa_stream_t stream);
/* joint attention over the L band copies of a window (192 tokens for L=3), images ordered (l, b):
 * S = scale*q.k^T + tables[l1*L+l2][rel_index] + (0/-100 intra|inter band mask) + shift mask; softmax; .v
 * ref: FrequencyWindowAttention.forward encoder_Uformer.py:256-310; kind 0 = intra, 1 = inter. hd = 28. */
int fa_joint_attn_fwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, float* o, int L, int B, int H, int W,
                      int heads, int hd, int shift, float scale, const float* tables, int kind, fa_stream_t stream);
int fa_joint_attn_bwd(const float* q, int64_t ldq, const float* kv, int64_t ldkv, const float* dout, float* dq,
                      float* dkv, int L, int B, int H, int W, int heads, int hd, int shift, float scale,
                      const float* tables, float* dtables, int kind, fa_stream_t stream);

/* ------------------------------------------------------------------ normalisation (K6, K9)
 * LayerNorm over the last dim (eps 1e-5). bwd: dx = dres + LN'(dy) (dres may be NULL); dgamma/dbeta ACCUMULATE.
 * ref: nn.LayerNorm at decoder_Uformer.py:666,744; encoder_Uformer.py:641,680,941. */
int fa_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                     int64_t rows, int C, fa_stream_t stream);
int fa_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                     const float* dres, float* dx, float* dgamma, float* dbeta, int64_t rows, int C, fa_stream_t stream);
/* BatchNorm2d (+LeakyReLU) on [B,C,S] (NCHW with S=H*W), the layout the encoder heads' raw reshape produces
 * (encoder_Uformer.py:978-982, encoder_ViT.py:195-199) and encoder_ResNet.py:9-16 uses.
 * stats: sums[c] = {sum, sumsq} as DOUBLES (2*C doubles), ACCUMULATED over (b,s); red likewise.  apply: y = lrelu(x*scale[c]+shift[c]) (y may be NULL)
 * and pooled[b][c] = mean_s y (pooled may be NULL).
 * bwd_reduce: with z = x*scale+shift, g = (dy ? dy : dpooled/S) * lrelu'(z): red[c] += {sum g, sum g*xhat}
 * bwd_apply: dx = scale[c] * (g - red0[c]/n - xhat * red1[c]/n)    (training)   or scale[c]*g (eval, red==NULL) */
int fa_bn_stats(const float* x, double* sums, int B, int C, int64_t S, fa_stream_t stream);
int fa_bn_apply(const float* x, const float* scale, const float* shift, float slope, float* y, float* pooled, int B,
                int C, int64_t S, fa_stream_t stream);
int fa_bn_bwd_reduce(const float* x, const float* mean, const float* rstd, const float* scale, const float* shift,
                     float slope, const float* dy, const float* dpooled, double* red, int B, int C, int64_t S,
                     fa_stream_t stream);
int fa_bn_bwd_apply(const float* x, const float* mean, const float* rstd, const float* scale, const float* shift,
                    float slope, const float* dy, const float* dpooled, const double* red, float* dx, int B, int C,
                    int64_t S, fa_stream_t stream);

/* The same BatchNorm on token layout [T, C] (NHWC; the ResNet encoder / DGRN path keeps NHWC throughout):
 * apply: y = lrelu(x*scale[c] + shift[c] + res) (slope 1.0 = no activation; res may be NULL).
 * bwd: g = dy * lrelu'(yout) (yout NULL: no activation); red[c] += {sum g, sum g*xhat} (doubles, caller zero-fills);
 *      dx = scale*(g - red0/T - xhat*red1/T) (training) or scale*g (training=0); dres = g if dres != NULL.
 * ref: nn.BatchNorm2d + LeakyReLU + residual add in ResBlock, encoder_ResNet.py:4-20. */
int fa_bn_tokens_stats(const float* x, double* sums, int64_t T, int C, fa_stream_t stream);
int fa_bn_tokens_apply(const float* x, const float* scale, const float* shift, const float* res, float slope, float* y,
                       int64_t T, int C, fa_stream_t stream);
int fa_bn_tokens_bwd(const float* x, const float* mean, const float* rstd, const float* scale, const float* yout,
                     float slope, const float* dy, double* red, float* dx, float* dres, int64_t T, int C, int training,
                     fa_stream_t stream);
/* AdaptiveAvgPool2d(1) on tokens (encoder_ResNet.py:30): out[b][c] = mean_t x[b][t][c]; bwd broadcasts dy/HW */
int fa_token_mean_fwd(const float* x, float* out, int B, int HW, int C, fa_stream_t stream);
int fa_token_mean_bwd(const float* dy, float* dx, int B, int HW, int C, fa_stream_t stream);

/* ------------------------------------------------------------------ convolutions on tokens (K4, K5)
 * depthwise 3x3 of LeFF in token layout with the second GELU fused (leff.py:85-86,100-112); the first GELU is the
 * producing GEMM's epilogue (which also stores the pre-activation u1 for the backward):
 *   fwd: u2 = dwconv(h1) + b ; h2 = gelu(u2)            (h1 = gelu(u1)).  u2_mode = 1 stores gelu'(u2) in `u2` instead
 *        (the only thing the backward needs u2 for: the dX contraction then multiplies by it with aux_act = 4, no
 *        transcendental in its epilogue); u2 or h2 may be NULL (inference stores h2 only).
 *   bwd: du1 = gelu'(u1) * dwconv^T(du2)  (u1 NULL: plain adjoint) ; dw, db ACCUMULATE (dw NULL: skipped).  h1 NULL with u1
 *        set: h1 = gelu(u1) is recomputed from the u1 values the kernel loads anyway (3.5 instead of 4.5 passes). */
int fa_dwconv3x3_fwd(const float* h1, const float* w, const float* b, float* u2, float* h2, int u2_mode, int B, int H,
                     int W, int C, fa_stream_t stream);
int fa_dwconv3x3_bwd(const float* du2, const float* h1, const float* u1, const float* w, float* du1, float* dw,
                     float* db, int B, int H, int W, int C, fa_stream_t stream);
/* patch gather / its adjoint for dense convs run as GEMMs: col[(b,oy,ox)][(ky,kx,ci)].
 * nchw_in: x is [B,C,H,W] instead of tokens.  ref: Conv2d at decoder_Uformer.py:418 (4x4 s2 p1), :457,:480 (3x3 s1 p1),
 * decoder_DGRN.py:5-6, encoder_ResNet.py:8-15. */
int fa_im2col(const float* x, float* col, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int nchw_in,
              fa_stream_t stream);
int fa_col2im(const float* col, float* dx, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
              fa_stream_t stream);
/* 3x3 s1 p1 convolution on NHWC tokens as an IMPLICIT GEMM: the patch matrix col[(b,y,x)][(ky,kx,ci)] is never written -
 * the TMA producer of the tcgen05 contraction addresses x [B,H,W,Cin] through a 4-D tensor map and lets the
 * out-of-bounds zero fill do the padding, so the layer reads x once instead of writing and re-reading 9x its size.
 *   fa_conv3x3_gemm : y[T, Cout] (row pitch ldy) = epi(col . wk^T), wk = [Cout][(ky,kx,ci)], epilogue as fa_gemm.
 *                     The data gradient is the same call on dY with the flipped, transposed weight
 *                     wk'[ci][((2-ky),(2-kx),co)] = wk[co][(ky,kx,ci)].
 *   fa_conv3x3_wgrad: dwk[Cout][(ky,kx,ci)] (+)= g[T, Cout]^T . col ; dbias (may be NULL) += column sums of g.
 * Eligible: Cin % 32 == 0, (H*W) % 128 == 0, W a multiple or a divisor of 128 (wgrad: W % 32 == 0); anything else is an
 * error (the caller keeps the explicit fa_im2col + fa_gemm path for those layers).  backend: 0 / 2 = 3xTF32, 4, 5.
 * ref: Conv2d at decoder_DGRN.py:5-6 (58 layers of DGRN), encoder_ResNet.py:8-15. */
int fa_conv3x3_gemm(const float* x, const float* wk, float* y, int B, int H, int W, int Cin, int Cout, int64_t ldy,
                    const FaGemmEpilogue* epi, int backend, fa_stream_t stream);
int fa_conv3x3_wgrad(const float* g, int64_t ldg, const float* x, float* dwk, int B, int H, int W, int Cin, int Cout,
                     int accumulate, float* dbias, int backend, fa_stream_t stream);
/* 3x3 s1 p1 conv from tokens [B,H*W,C] to a FEW output channels, written as an NCHW image [B,Co,H*W] with the bias and
 * an optional residual image added: OutputProj + the network's global skip (decoder_Uformer.py:476-499, :1171).
 * C % 4 == 0, C <= 128, Co <= 4.  wk = [Co][(ky,kx,ci)].  No patch matrix is built (it would be 9*C/Co times the output).
 * bwd: dt [B,H*W,C] (may be NULL), dW [Co][(ky,kx,ci)] and db [Co] ACCUMULATE (caller zero-fills; either may be NULL). */
int fa_conv3x3_out_fwd(const float* t, const float* wk, const float* bias, const float* ximg, float* out, int B, int H,
                       int W, int C, int Co, fa_stream_t stream);
int fa_conv3x3_out_bwd(const float* t, const float* wk, const float* dout, float* dt, float* dW, float* db, int B, int H,
                       int W, int C, int Co, fa_stream_t stream);
/* ConvTranspose2d k2 s2 scatter (decoder_Uformer.py:438): y[b][2y+ky][2x+kx][co] = g[(b,y,x)][(ky,kx,co)];
 * y has row stride ldy (writes the left half of the skip-concat buffer, :1162). bwd is the gather. */
int fa_pixel_shuffle2_fwd(const float* g, float* y, int64_t ldy, int B, int H, int W, int Co, fa_stream_t stream);
int fa_pixel_shuffle2_bwd(const float* dy, int64_t ldy, float* dg, int B, int H, int W, int Co, fa_stream_t stream);
/* strided 2-D copy / layout helpers */
int fa_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int cols, fa_stream_t stream);
int fa_add2d(const float* a, int64_t lda, const float* b, int64_t ldb, float* dst, int64_t ldd, int64_t rows, int cols,
             fa_stream_t stream);
/* dst[r][c] = src[r][c] * rowscale[r / rows_per_scale]  (DropPath backward: per-sample scale of a gradient) */
int fa_scale_rows(const float* src, const float* rowscale, int rows_per_scale, float* dst, int64_t rows, int cols,
                  fa_stream_t stream);
/* y[b][c][hw] = t[b][hw][c] (+ res[b][c][hw]); and the reverse */
int fa_tokens_to_nchw(const float* t, const float* res, float* y, int B, int HW, int C, fa_stream_t stream);
int fa_nchw_to_tokens(const float* x, float* t, int B, int HW, int C, fa_stream_t stream);

/* ------------------------------------------------------------------ DGRN (K7, K8)
 * DCNv2 3x3 s1 p1, groups 1, deformable_groups 1 on NHWC tokens.  om [B*H*W][ldom >= 27]: raw output of conv_offset_mask
 * in the reference's channel order (deform_conv.py:59-62): o1 = ch 0..8, o2 = ch 9..17, mask = sigmoid(ch 18..26);
 * offset = cat(o1,o2) so tap k reads (dy,dx) = (om[2k], om[2k+1]) of the 18 offset channels.
 * fwd writes col [B*H*W][9*C] = mask_k * bilinear(x, p + p_k + offset_k) for the contraction with weight (fa_gemm).
 * bwd: from dcol -> dx (atomic scatter, caller zero-fills), dom [B*H*W][ldom] (entries 0..26 of every row written).
 * ldom = 32 lets the offset / mask convolution run as the implicit GEMM (fa_conv3x3_gemm wants Cout % 4 == 0) and its
 * backward take dom as a 32-channel token tensor.
 * ref: DCN_layer.forward deform_conv.py:56-67 + mmcv modulated_deform_conv2d (absent; parity unpinned). */
int fa_dcn_im2col(const float* x, const float* om, int ldom, float* col, int B, int H, int W, int C, fa_stream_t stream);
int fa_dcn_col2im(const float* x, const float* om, int ldom, const float* dcol, float* dx, float* dom, int B, int H, int W,
                  int C, fa_stream_t stream);
/* SFT + DGM tail (decoder_DGRN.py:22-32,49-57,79-81): out = act(x + dcn + x*gamma + beta) */
int fa_sft_fuse_fwd(const float* x, const float* dcn, const float* gamma, const float* beta, float* out, int64_t n,
                    float slope, fa_stream_t stream);
int fa_sft_fuse_bwd(const float* x, const float* dcn, const float* gamma, const float* beta, const float* dout,
                    float* dx, float* ddcn, float* dgamma, float* dbeta, int64_t n, float slope, fa_stream_t stream);

/* ------------------------------------------------------------------ band-weight (lambda) predictor (a10)
 * out[b][h] (at out + b*ld_b + h*ld_h) = W2 lrelu(W0 (fc_w (stats[b] * ln_w + ln_b) + fc_b) + b0, 0.1) + b2 with
 * stats [B, D] = the token mean of the affine-free LayerNorm of one band's encoder features.
 * params / grads: 8 HOST-side arrays of device pointers {ln_w[D], ln_b[D], fc_w[heads,D], fc_b[heads], w0[heads,heads],
 * b0[heads], w2[heads,heads], b2[heads]}.  bwd ACCUMULATES into grads[i] and dstats [B, D] (dstats may be NULL).
 * ref: WindowAttention.__init__ / forward, decoder_Uformer.py:178-193,280-284 (mlp_head[i], avg[i], mlp[i]). */
int fa_band_coef_fwd(const float* stats, const float* const* params, float* out, int B, int D, int heads, int64_t ld_b,
                     int64_t ld_h, fa_stream_t stream);
int fa_band_coef_bwd(const float* stats, const float* const* params, const float* dout, int64_t ld_b, int64_t ld_h,
                     float* dstats, float* const* grads, int B, int D, int heads, fa_stream_t stream);

/* ------------------------------------------------------------------ elementwise / loss / optimiser (K10, K11)
 * y = act(x) and dx = dy * act'(x) */
int fa_act_fwd(const float* x, float* y, int64_t n, int act, float p, fa_stream_t stream);
int fa_act_bwd(const float* dy, const float* x, float* dx, int64_t n, int act, float p, fa_stream_t stream);
/* loss[0] += mean|a-b| ; grad = sign(a-b) * gscale / n   (nn.L1Loss, train.py:65,89) */
int fa_l1_loss(const float* a, const float* b, float* loss, float* grad, int64_t n, float gscale, fa_stream_t stream);
/* Training-pair generator (utils/dataset_utils.py:97-135 pixel work, whole batch per launch): for sample i,
 * meta[i] = {clean_offset, degraded_offset or -1, H, W, y0, x0, mode, seed} (int64; offsets into the uint8 HWC image
 * pool; W = row length of the image).  Cuts the P x P patch at (y0, x0) from the clean and the degraded image, applies
 * augmentation `mode` (0..7 = image_utils.data_augmentation) and ToTensor: out_* [B,3,P,P] float in [0,1].
 * degraded_offset = -1 synthesises the degradation as the reference does for 'denoising_*':
 *   clip(clean + n * sigma[i], 0, 255).astype(uint8), n ~ N(0,1) either read from noise [B,P,P,3] (patch coordinates
 *   BEFORE augmentation) or, when noise is NULL, drawn from a stateless generator keyed on (seed, absolute pixel,
 *   channel) so that the two crops of a pair share the noise of their overlap, as one noisy image would give. */
int fa_crop_augment(const uint8_t* pool, const int64_t* meta, const float* sigma, const float* noise, float* out_degraded,
                    float* out_clean, int B, int P, fa_stream_t stream);
/* k = k*m + q*(1-m) over a flat buffer (moco.py:45-50) */
int fa_momentum_update(float* k, const float* q, int64_t n, float m, fa_stream_t stream);
/* torch.optim.Adam step over a flat buffer (train.py:63,96); step >= 1 */
int fa_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                 int step, float grad_scale, fa_stream_t stream);
/* the same update with the step-dependent quantities held in DEVICE memory: state = {float lr, int32 step count}.
 * fa_adam_tick does count += 1 (one thread); fa_adam_step_state reads lr and the count at run time and derives
 * lr / (1 - beta1^t) and 1 / sqrt(1 - beta2^t) inside the kernel (double precision, as torch's host code).  A captured
 * CUDA graph of the train step is therefore self-contained: replays never read host memory, however far the host runs
 * ahead of the device, and a scheduled lr is one stream-ordered device write. */
int fa_adam_tick(float* state, fa_stream_t stream);
int fa_adam_step_state(float* p, const float* g, float* m, float* v, int64_t n, const float* state, float beta1,
                       float beta2, float eps, float grad_scale, fa_stream_t stream);
/* dst[i] = src[i] rounded to the nearest TF32 (ties away from zero; low 13 mantissa bits zero): the pre-rounded weight
 * operand of the 1xTF32 / 2xTF32 contractions (FaGemmEpilogue.b_is_tf32); the train step refreshes one rounded copy of
 * each flat parameter buffer per step */
int fa_round_tf32(const float* src, float* dst, int64_t n, fa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FREQAIR_H */
