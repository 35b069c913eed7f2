/* unetca_b200.h — C ABI of libunetca_b200.so: the B200-native U-Net-CA hot path.
 *
 * The reference (Createroner/InSAR-Unet-CA) has no FFI; its hot path is reached through the torch.nn.Module
 * protocol (Unet-ChannalAttention.py, "UCA" below).  This ABI is what a drop-in nn.Module binds instead of the
 * ATen ops that UCA's modules dispatch to; each entry cites the reference call site it replaces.
 *
 * Rules for every function:
 *   - plain pointers and sizes only; every pointer is a CUDA *device* pointer unless stated otherwise;
 *   - the caller owns all memory (tensors, scratch); the library allocates nothing.  Its only state is
 *     process-global and read-only on the data path: cached device properties (SM count, kernel attributes), the
 *     implementation switch below, and the development knobs of unetca_b200_tuning.h (A/B sweeps and cross-checks
 *     only — not part of the drop-in ABI, not re-entrant; the product path never calls them);
 *   - stream-ordered on `stream` (a cudaStream_t), no host synchronisation, no exceptions;
 *   - returns 0 (or a documented non-negative count) on success, <0 on error; unetca_last_error() returns the
 *     thread-local message.  There is no CPU fallback: without a CUDA device every op fails.
 *   - activations are NHWC: element (b,h,w,c) at ((b*H+h)*W+w)*ld + c with ld >= C (a tensor may be a channel
 *     slice of a concat buffer); `dtype` selects their storage: UNETCA_DTYPE_F32 (fp32 parity mode, FFMA
 *     contractions) or UNETCA_DTYPE_BF16 (tcgen05 contractions, fp32 accumulate).  Parameters, statistics,
 *     gradients of parameters, logits and the loss are always fp32.
 *   - `parts` arguments are fp32 scratch for deterministic two-stage reductions; they must hold
 *     unetca_max_parts(B) rows of the documented width.
 *
 * This header is parsed by the Python binding (one declaration per statement, `name(type arg, ...)`).
 */
#ifndef UNETCA_B200_H
#define UNETCA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNETCA_DTYPE_F32 0
#define UNETCA_DTYPE_BF16 1

/* ---- runtime ------------------------------------------------------------------------------------------- */
const char* unetca_last_error(void);
int unetca_abi_version(void);
int unetca_num_sms(void);
int unetca_max_parts(int B);
/* 0 (default): bf16 -> tcgen05 kernels, fp32 -> FFMA kernels; 1: FFMA kernels for both (cross-check only) */
void unetca_set_conv_impl(int impl);
int unetca_get_conv_impl(void);

/* ---- module boundary: layout and parameter packing ------------------------------------------------------ */
/* network input (B,Cin,H,W) NCHW fp32 (UCA:343 `model(images)`) -> im2col rows [B*H*W][Kpad], k = tap*Cin + c */
int unetca_im2col3x3_nchw(int dtype, const float* x, void* col, int B, int Cin, int H, int W, int Kpad, void* stream);
int unetca_nchw_to_nhwc(int dtype, const float* x, void* y, int ld, int B, int C, int H, int W, void* stream);
int unetca_nhwc_to_nchw(int dtype, const void* x, int ld, float* y, int B, int C, int H, int W, void* stream);
/* device side of the reference's input preprocessing (UCA:200-210, 428-433): uint8 tile -> ToTensor (/255) ->
 * Normalize(mean, std) fp32; uint8 mask -> ToTensor().long() int64 (255 -> 1, everything else -> 0) */
int unetca_prep_u8(const uint8_t* img, const uint8_t* mask, float* out, long long* lab, long n, float mean, float stdv, void* stream);
/* nn.Conv2d weight (O,C,3,3) -> wf [O][ldk] (k = tap*C + c, zero padded) and optional dgrad operand wd [C][9*O] */
int unetca_pack_conv3x3_weight(int dtype, const float* w, void* wf, int ldk, void* wd, int O, int C, void* stream);
/* nn.ConvTranspose2d weight (Cin,Cout,2,2) -> wf [4*Cout][Cin] and wd [Cin][4*Cout] */
int unetca_pack_convT_weight(int dtype, const float* w, void* wf, void* wd, int Cin, int Cout, void* stream);

/* ---- contractions (tensor pipe) -------------------------------------------------------------------------- */
/* nn.Conv2d(k=3,pad=1) forward without bias, UCA:81,84; also its dgrad (pass dy, wd and swap C/O).
 * stat_parts (optional) receives partial per-channel sum / sum-of-squares of y: [*nparts][2][O]. */
int unetca_conv3x3_fwd(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W, int C, int O, float* stat_parts, int* nparts, void* stream);
/* the same convolution for narrow outputs (O % 64 == 0, H even) through the row-pair layout of the tcgen05 path
 * (csrc/conv_tc.cu, tc_conv3x3_pixn_kernel); w_pair [2*O][12*C] = unetca_pack_conv3x3_pair(w [O][ld]).  bf16 only. */
int unetca_conv3x3_fwd_paired(int dtype, const void* x, int ldx, const void* w_pair, void* y, int ldy, int B, int H, int W, int C, int O, float* stat_parts, int* nparts, void* stream);
/* the 64 -> 64 channel layers (inc / conv4 second convs and their dgrads, UCA:84 at full resolution), H even: row-pair
 * layout with the whole filter resident in shared memory and haloed lattice tiles (tc_conv3x3_rp64_kernel); w is the
 * ordinary packed filter [64][ldk >= 576] of unetca_pack_conv3x3_weight.  bf16 only. */
/* First conv of a decoder block: its input is torch.cat([skip, up], dim=1) (UCA:140,146,152,158).  The two halves stay two
 * DENSE tensors (x: channels [0, C1) = skip, x2: [C1, C) = upsampled; C1 % 64 == 0): the concat costs nothing and nobody writes or
 * reads half-pixels at a doubled stride.  fwd_cat: O % 128 == 0 with w = packed filter [O][9*C], or O == 64 && C == 128 with
 * w = kw-stacked filter; scale/shift non-null: eval-mode BatchNorm + ReLU in the epilogue (as unetca_conv3x3_bnrelu_fwd), sq_parts:
 * SE squeeze sums (O % 128 == 0 only).  wgrad_cat: as unetca_conv3x3_wgrad.  bf16 only; other shapes: UNETCA_ERR_UNSUPPORTED. */
int unetca_conv3x3_fwd_cat(int dtype, const void* x, int ldx, const void* x2, int ldx2, int C1, const void* w, void* y, int ldy, int B, int H, int W, int C, int O, float* stat_parts, const float* scale, const float* shift, float* sq_parts, int* nparts, void* stream);
int unetca_conv3x3_wgrad_cat(int dtype, const void* dy, int lddy, const void* x, int ldx, const void* x2, int ldx2, int C1, float* ws, long ws_floats, int B, int H, int W, int C, int O, float* dw, void* stream);
/* unetca_conv3x3_fwd (O % 128 == 0) writing output channels [0, split) to y and [split, O) to y2: the dgrad of a decoder block's
 * first conv (its input is torch.cat([skip, up]), UCA:140) leaves d(skip) and d(up) as two dense tensors.  bf16 only. */
int unetca_conv3x3_fwd_split(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, void* y2, int ldy2, int split, int B, int H, int W, int C, int O, float* stat_parts, int* nparts, void* stream);
int unetca_conv3x3_fwd_rp64(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W, float* stat_parts, int* nparts, void* stream);
/* dgrad of a block's second conv (dy -> dA1, wd = dgrad-packed filter [O][ldk]) that also leaves the statistics of the
 * ReLU + BatchNorm backward that follows (autograd of UCA:82-83): parts [*nparts][2][O] = (sum dz, sum dz*(y1 - mean)) with
 * dz = dA1 * (scale*y1 + shift > 0), y1 the saved conv output — the input of unetca_bn_bwd_finalize, without the
 * unetca_bn_bwd_reduce pass over dA1 and y1.  O % 128 == 0, or C = O = 64 with an even H; else UNETCA_ERR_UNSUPPORTED (-3). */
int unetca_conv3x3_dgrad_bnstats(int dtype, const void* dy, int lddy, const void* wd, int ldk, void* da, int ldda, int B, int H, int W, int C, int O, const void* y1, int ldy1, const float* scale, const float* shift, const float* mean, float* parts, int* nparts, void* stream);
int unetca_pack_conv3x3_pair(int dtype, const void* w, int ld, void* w_pair, int rows, int C, void* stream);
/* the same convolution for exactly 64 output channels and C = 64 or 128 through the kw-stacked layout of the tcgen05 path
 * (csrc/conv_tc.cu, tc_conv3x3_kw_kernel: N = 3 kw taps x 64 channels, horizontal shift-add in the epilogue, filter
 * resident in shared memory); w_kw [9*C][64] = unetca_pack_conv3x3_kw(w [64][ld]).  bf16 only, any H, W. */
int unetca_conv3x3_fwd_kw(int dtype, const void* x, int ldx, const void* w_kw, void* y, int ldy, int B, int H, int W, int C, float* stat_parts, int* nparts, void* stream);
int unetca_pack_conv3x3_kw(int dtype, const void* w, int ld, void* w_kw, int C, void* stream);
/* inference forms (bf16 tcgen05 path): conv3x3 + eval-mode BatchNorm folded to (scale, shift) by unetca_bn_fold_eval + ReLU
 * applied in the conv epilogue — UCA:81-86 under model.eval() (UCA:276) without writing the pre-activation tensor.
 * layout 0: w = packed filter [O][9*C], O % 128 == 0; 1: w = pair-packed filter, H even; 2: w = kw-stacked filter;
 * 3: w = packed filter, C = O = 64, H even (resident-filter row-pair kernel).
 * sq_parts (optional, layouts 0/1; zero-filled by the caller: B * unetca_num_sms() * O floats) receives the SE squeeze as
 * per-image partial channel sums [B][*nparts][O] of the stored activation, ready for unetca_se_fc */
int unetca_conv3x3_bnrelu_fwd(int dtype, const void* x, int ldx, const void* w, int layout, void* y, int ldy, int B, int H, int W, int C, int O, const float* scale, const float* shift, float* sq_parts, int* nparts, void* stream);
int unetca_first_pairs_bnrelu_fwd(int dtype, const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O, const float* scale, const float* shift, void* stream);
/* first conv (Cin <= 5, H even) in the row-pair layout of the tcgen05 path (bf16): one im2col row per pixel PAIR
 * (rows 2i, 2i+1 of a column) holding their shared 4x3 patch, colp [B*(H/2)*W][64]; pair-packed filter wp [2*O][64];
 * forward = one GEMM with 128x256x16 MMAs (+ BatchNorm partial sums), weight gradient from the same colp */
int unetca_im2col_pairs(int dtype, const float* x, void* colp, int B, int Cin, int H, int W, void* stream);
int unetca_pack_first_pairs(int dtype, const float* w, void* wp, int O, int Cin, void* stream);
int unetca_first_pairs_fwd(int dtype, const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O, float* stat_parts, int* nparts, void* stream);
int unetca_first_pairs_wgrad(int dtype, const void* dy, int lddy, const void* colp, float* ws, long ws_floats, int B, int H, int W, int Cin, int O, float* dw, void* stream);
int unetca_first_pairs_fold(const float* ws, int nsplit, int O, int Cin, float* dw, void* stream);
/* first conv (K = 9*Cin): out[m][n] = sum_k A[m][k] * Bw[n][k] over im2col rows */
int unetca_gemm_nt(int dtype, const void* A, int lda, const void* Bw, int ldb, void* out, int ldo, long M, int N, int K, float* stat_parts, int* nparts, void* stream);
/* weight gradient of conv3x3 -> dw (O,C,3,3) fp32; ws: split-K scratch (ws_floats floats) */
int unetca_conv3x3_wgrad(int dtype, const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats, int B, int H, int W, int C, int O, float* dw, void* stream);
int unetca_im2col_wgrad(int dtype, const void* dy, int lddy, const void* col, int Kpad, float* ws, long ws_floats, long npix, int Cin, int O, float* dw, void* stream);
/* nn.ConvTranspose2d(k=2,s=2), UCA:112,115,118,121 (+bias); out may be the upper half of a concat buffer */
int unetca_convT2x2_fwd(int dtype, const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_convT2x2_dgrad(int dtype, const void* dout, int ldd, const void* wdg, void* dx, int ldx, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_convT2x2_wgrad(int dtype, const void* x, int ldx, const void* dout, int ldd, float* ws, long ws_floats, int B, int h, int wd, int Cin, int Cout, float* dw, void* stream);

/* ---- BatchNorm2d + ReLU, UCA:82-83,85-86 ------------------------------------------------------------------ */
int unetca_chan_stats(int dtype, const void* y, int ld, int C, long npix, float* parts, int* nparts, void* stream);
/* train: batch statistics (biased var to normalise, unbiased into running_var, momentum), fused affine
 * scale = gamma*invstd, shift = beta - mean*scale; conv_bias only shifts running_mean (BN cancels it) */
int unetca_bn_finalize_train(const float* parts, int nparts, int C, long count, const float* conv_bias, const float* gamma, const float* beta, float* running_mean, float* running_var, float momentum, float eps, float* mean, float* invstd, float* scale, float* shift, void* stream);
/* eval (UCA:276): scale = gamma/sqrt(running_var+eps), shift = beta + (conv_bias - running_mean)*scale */
int unetca_bn_fold_eval(int C, const float* conv_bias, const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps, float* scale, float* shift, void* stream);
/* out = relu(scale*y+shift) and/or per-image channel sums of it (SE squeeze): pool_parts [B][*nparts][C] */
int unetca_bn_relu(int dtype, const void* y, int ldy, void* out, int ldo, int B, long pix_per_img, int C, const float* scale, const float* shift, float* pool_parts, int* nparts, void* stream);
int unetca_bn_bwd_reduce(int dtype, const void* dout, int ldd, const void* y, int ldy, int B, long pix_per_img, int C, const float* scale, const float* shift, const float* mean, const float* invstd, const float* s, const float* dp, float* parts, int* nparts, void* stream);
int unetca_bn_bwd_finalize(const float* parts, int nparts, int C, long count, const float* gamma, const float* invstd, float* dgamma, float* dbeta, float* coef, void* stream);
int unetca_bn_bwd_apply(int dtype, const void* dout, int ldd, const void* y, int ldy, void* dy, int lddy, int B, long pix_per_img, int C, const float* scale, const float* shift, const float* mean, const float* invstd, const float* s, const float* dp, const float* coef, void* stream);

/* ---- SELayer, UCA:45-72, and MaxPool2d(2), UCA:106-109 ----------------------------------------------------- */
/* p = mean_hw, z = relu(W1 p), s = sigmoid(W2 z) from the squeeze partial sums */
int unetca_se_fc(const float* pool_parts, int nparts, int B, int C, int Cr, long hw, const float* w1, const float* w2, float* p, float* z, float* s, void* stream);
/* SELayer.forward called on its own (UCA:61-72; NCHW fp32, any sign): plane-wise passes around unetca_se_fc /
 * unetca_se_fc_bwd.  plane_dot: out[plane] = sum a * (b ? b : 1) over the hw contiguous floats of each (b,c) plane (the
 * squeeze, and ds = sum dy * x); plane_scale_add: out = x * s[plane] + (t ? t[plane] * tscale : 0) (y = x * s, and
 * dx = dy * s + dp / HW) */
int unetca_plane_dot(const float* a, const float* b, long nplanes, long hw, float* out, void* stream);
int unetca_plane_scale_add(const float* x, const float* s, const float* t, float tscale, long nplanes, long hw, float* out, void* stream);
/* out = relu(scale*y+shift) * s[b,c] (s null: no SE); pooled/pos non-null: fused 2x2 max-pool of out with
 * torch's first-max / NaN rule, pos = 1-byte window position dh*2+dw */
int unetca_se_scale_pool(int dtype, const void* y, int ldy, void* out, int ldo, void* pooled, int ldp, uint8_t* pos, int B, int H, int W, int C, const float* scale, const float* shift, const float* s, void* stream);
/* standalone pool; idx64 (optional) = torch's (B,C,H/2,W/2) int64 flat indices h*W+w */
int unetca_maxpool2x2(int dtype, const void* x, int ldx, void* pooled, int ldp, uint8_t* pos, long long* idx64, int B, int H, int W, int C, void* stream);
int unetca_pool_bwd_add(int dtype, const void* skip_grad, int lds, const void* dpooled, int ldp, const uint8_t* pos, void* dx, int ldx, int B, int H, int W, int C, void* stream);
/* bilinear resize guard of the decoder (UCA:138-157: F_T.resize(x, skip.shape, BILINEAR) when H or W is not in 16*N):
 * (B,h,w,C) -> (B,H,W,C), align_corners=False, and its adjoint */
int unetca_resize_bilinear_fwd(int dtype, const void* x, int ldx, int h, int w, void* out, int ldo, int H, int W, int B, int C, void* stream);
int unetca_resize_bilinear_bwd(int dtype, const void* dout, int ldd, int H, int W, void* dx, int ldx, int h, int w, int B, int C, void* stream);
int unetca_se_bwd_reduce(int dtype, const void* dout, int ldd, const void* y, int ldy, int B, long pix_per_img, int C, const float* scale, const float* shift, float* parts, int* nparts, void* stream);
int unetca_se_fc_bwd(const float* parts, int nparts, int B, int C, int Cr, const float* w1, const float* w2, const float* p, const float* z, const float* s, float* dpre2, float* dz, float* dp, float* dw1, float* dw2, void* stream);
/* SE squeeze as two per-image partial sums (count of active pixels, masked sum of y): parts [B * *nparts][2][C];
 * se_fc3 derives p = mean_hw relu(a*y+b) from them, runs the FC chain, and keeps (sum m, sum m*(y-mean)) for the backward */
int unetca_se_squeeze(int dtype, const void* y, int ldy, int B, long pix_per_img, int C, const float* scale, const float* shift, float* parts, int* nparts, void* stream);
int unetca_se_fc3(const float* parts2, int nparts, int B, int C, int Cr, long hw, const float* w1, const float* w2, const float* scale, const float* shift, const float* mean, float* p, float* z, float* s, float* sums34, void* stream);
/* merged SE + ReLU + BN backward reduction (one pass over dO, Y2): parts [B * *nparts][2][C]; FC chain; BN finalize */
int unetca_se_bn_bwd_reduce(int dtype, const void* dout, int ldd, const void* y, int ldy, int B, long pix_per_img, int C, const float* scale, const float* shift, const float* mean, float* parts, int* nparts, void* stream);
int unetca_se_fc_bwd_fused(const float* parts, int nparts, int B, int C, int Cr, const float* w1, const float* w2, const float* p, const float* z, const float* s, const float* scale, const float* shift, const float* mean, const float* sums34, float* sums, float* dpre2, float* dz, float* dp, float* dw1, float* dw2, void* stream);
int unetca_bn_bwd_finalize_se(const float* sums, int B, int C, long count, long pix_per_img, const float* gamma, const float* invstd, const float* s, const float* dp, float* dgamma, float* dbeta, float* coef, void* stream);
/* the same two passes for an ENCODER block, with the gradient of the block output rebuilt on the fly from the skip
 * gradient sg (full resolution) and the pooled gradient routed through the max-pool positions (no pool_bwd_add pass) */
int unetca_se_bn_bwd_reduce_pool(int dtype, const void* sg, int lds, const void* dpooled, int ldp, const uint8_t* pos, const void* y, int ldy, int B, int H, int W, int C, const float* scale, const float* shift, const float* mean, float* parts, int* nparts, void* stream);
int unetca_bn_bwd_apply_pool(int dtype, const void* sg, int lds, const void* dpooled, int ldp, const uint8_t* pos, const void* y, int ldy, void* dy, int lddy, int B, int H, int W, int C, const float* scale, const float* shift, const float* mean, const float* invstd, const float* s, const float* dp, const float* coef, void* stream);
int unetca_chan_sum(int dtype, const void* x, int ld, int C, long npix, float* parts, float* out, void* stream);
/* out[i] = sum over nrows rows of parts[row*row_stride + i] (i < n): second stage for per-CTA channel sums, e.g. the
 * ConvTranspose2d bias gradient (UCA:114) taken from the statistics epilogue of the convolution that writes dcat */
int unetca_sum_rows(const float* parts, int nrows, long row_stride, int n, float* out, void* stream);

/* ---- outc 1x1 conv -> class logits (UCA:125,162), CrossEntropyLoss(ignore_index) (UCA:465,344), argmax (UCA:220) */
int unetca_outc_fwd(int dtype, const void* x, int ldx, int C, const float* w, const float* bias, int nc, float* logits, int B, long HW, void* stream);
/* The output head fused with the last DoubleConv's elementwise passes (UCA:86-94 of conv4, then UCA:162; bf16, dense C = 64,
 * nc <= 2, HW % 128 == 0 — otherwise UNETCA_ERR_UNSUPPORTED (-3) and the caller runs the separate passes).  The block output
 * h = relu(scale*y2 + shift) * s[b,c] and its gradient W^T dlogits are per-pixel functions of y2 and of the 8 bytes of dlogits:
 * neither full-resolution tensor is written or read.
 *   se_scale_outc_fwd:   y2 -> logits (B,nc,HW) NCHW fp32
 *   outc_bn_bwd_reduce:  (g = dlogits up to *gscale, y2) -> parts_bn rows [B * *nparts][2][C] exactly as unetca_se_bn_bwd_reduce /
 *                        unetca_bn_bwd_reduce leave them (*nparts rows per image), and the outc gradients dw (nc,C), db (nc);
 *                        s: the SE scale of the block (null without SE); parts_oc: scratch, parts_oc_floats floats
 *   outc_bn_bwd_apply:   (g, y2) -> dY2, arguments as unetca_bn_bwd_apply */
int unetca_se_scale_outc_fwd(int dtype, const void* y, int ldy, int B, long HW, int C, const float* scale, const float* shift, const float* s, const float* w, const float* bias, int nc, float* logits, void* stream);
int unetca_outc_bn_bwd_reduce(int dtype, const float* g, const float* gscale, const float* w, int nc, const void* y, int ldy, int B, long HW, int C, const float* scale, const float* shift, const float* mean, const float* s, float* parts_bn, int* nparts, float* parts_oc, long parts_oc_floats, float* dw, float* db, void* stream);
int unetca_outc_bn_bwd_apply(int dtype, const float* g, const float* gscale, const float* w, int nc, const void* y, int ldy, void* dy, int lddy, int B, long HW, int C, const float* scale, const float* shift, const float* mean, const float* invstd, const float* s, const float* dp, const float* coef, void* stream);
int unetca_outc_bwd(int dtype, const float* g, const float* gscale, const void* x, int ldx, void* dx, int lddx, int C, const float* w, int nc, int B, long HW, float* parts, float* dw, float* db, void* stream);
/* loss_out[0] = mean CE over valid pixels (NaN when none), loss_out[1] = #valid; g = un-normalised dlogits;
 * gscale_out[0] = upstream/#valid; mask = argmax class map (first maximum wins); target/g/mask optional */
int unetca_cross_entropy(const float* logits, const long long* target, int nc, int B, long HW, long long ignore_index, const float* upstream, float* g, long long* mask, float* parts, float* loss_out, float* gscale_out, void* stream);
/* compute_metrics (UCA:214-269) on the device: counts[(nc+1)][nc] int64, row = label (row nc = label values outside
 * [0,nc) other than ignore_index), column = argmax class of the logits (first maximum wins, UCA:220); pixels with
 * label == ignore_index are dropped (UCA:223).  parts: unetca_max_parts(B) * (nc+1)*nc 8-byte words of scratch */
int unetca_confusion_counts(const float* logits, const long long* target, int nc, int B, long HW, long long ignore_index, void* parts, long long* counts, void* stream);

/* ---- implementation-specific contraction entry points (exported for the cross-check tests) ----------------- */
int unetca_tc_conv3x3_fwd(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W, int C, int O, float* stat_parts, void* stream);
int unetca_tc_conv3x3_fwd_paired(const void* x, int ldx, const void* w_pair, void* y, int ldy, int B, int H, int W, int C, int O, float* stat_parts, void* stream);
int unetca_tc_pack_pair(const void* w, int ld, void* w_pair, int rows, int C, void* stream);
int unetca_tc_conv3x3_fwd_kw(const void* x, int ldx, const void* w_kw, void* y, int ldy, int B, int H, int W, int C, float* stat_parts, void* stream);
int unetca_tc_pack_kw(const void* w, int ld, void* w_kw, int C, void* stream);
int unetca_tc_conv3x3_bnrelu_fwd(const void* x, int ldx, const void* w, int layout, void* y, int ldy, int B, int H, int W, int C, int O, const float* scale, const float* shift, float* sq_parts, void* stream);
int unetca_tc_first_pairs_bnrelu_fwd(const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O, const float* scale, const float* shift, void* stream);
int unetca_tc_first_pairs_fwd(const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O, float* stat_parts, void* stream);
int unetca_tc_first_pairs_wgrad(const void* dy, int lddy, const void* colp, float* ws, long ws_floats, int B, int H, int W, int O, void* stream);
int unetca_tc_gemm_nt(const void* A, int lda, const void* Bw, int ldb, void* out, int ldo, long M, int N, int K, float* stat_parts, void* stream);
int unetca_tc_convT_fwd(const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_tc_convT_dgrad(const void* dout, int ldd, const void* wdg, void* dx, int ldx, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_tc_conv3x3_wgrad(const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats, int B, int H, int W, int C, int O, void* stream);
int unetca_tc_conv3x3_wgrad_generic(const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats, int B, int H, int W, int C, int O, void* stream);
int unetca_tc_gemm_tn(const void* A, int lda, const void* Bm, int ldb, float* ws, long ws_floats, int M, int N, long K, void* stream);
int unetca_tc_convT_wgrad(const void* x, int ldx, const void* dout, int ldd, float* ws, long ws_floats, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_simt_conv3x3_fwd(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W, int C, int O, void* stream);
int unetca_simt_gemm_nt(int dtype, const void* A, int lda, const void* Bm, int ldb, void* out, int ldo, int M, int N, int K, void* stream);
int unetca_simt_conv3x3_wgrad(int dtype, const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats, int B, int H, int W, int C, int O, void* stream);
int unetca_simt_gemm_tn(int dtype, const void* A, int lda, const void* Bm, int ldb, float* ws, long ws_floats, int M, int N, long K, void* stream);
int unetca_simt_convT_fwd(int dtype, const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_simt_convT_dgrad(int dtype, const void* dout, int ldd, const void* w, void* dx, int ldx, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_simt_convT_wgrad(int dtype, const void* x, int ldx, const void* dout, int ldd, float* ws, long ws_floats, int B, int h, int wd, int Cin, int Cout, void* stream);
int unetca_wgrad_reduce(const float* ws, int nsplit, long split_stride, int mode, int D0, int D1, int ldn, float* dw, void* stream);

/* ---- optimizer step: optim.Adam(model.parameters(), lr) UCA:466, optimizer.step() UCA:346 ----
 * Multi-tensor Adam on the fp32 master weights (torch.optim.Adam arithmetic, no amsgrad / maximize).
 * hyper: 8 floats on the device, zero-initialised by the caller; hyper[0] is the step count.  unetca_adam_tick
 * advances it and derives the bias corrections on the device, so a captured train step replays correctly.
 * unetca_adam_step: table = host array of ntensors rows {p, g, m, v, numel} (64-bit each; fp32 device pointers).
 * unetca_adam_step_conv3x3: table = host array of nlayers rows {p, g, m, v, wf, wd, O, C}: (O,C,3,3) filters with O, C
 * multiples of 32; the stepped weights are also emitted as the packed operand copies of unetca_pack_conv3x3_weight
 * (wf [O][9*C] forward, wd [C][9*O] dgrad, storage type dtype; either may be 0).  Both return the launch count. */
int unetca_adam_tick(float* hyper, double lr, double beta1, double beta2, double eps, double weight_decay, void* stream);
int unetca_adam_step(const long long* table, int ntensors, const float* hyper, void* stream);
int unetca_adam_step_conv3x3(int dtype, const long long* table, int nlayers, const float* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif
