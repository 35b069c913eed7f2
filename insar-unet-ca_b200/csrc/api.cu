// api.cu — the contraction entry points of the C ABI (see include/unetca_b200.h).
//
// Each op picks its implementation from the storage type: bf16 -> tcgen05/TMEM/TMA kernels (conv_tc.cu),
// fp32 -> CUDA-core FFMA parity kernels (gemm_simt.cu).  unetca_set_conv_impl(1) forces the FFMA kernels for
// bf16 as well; the GPU tests use that to cross-check the tensor-core path on identical bf16 operands.
// There is no CPU path: every op needs a CUDA device and fails with an error string otherwise.
#include "common.cuh"

extern "C" {
int unetca_tc_conv3x3_fwd(const void*, int, const void*, int, void*, int, int, int, int, int, int, float*, void*);
int unetca_tc_conv3x3_fwd_paired(const void*, int, const void*, void*, int, int, int, int, int, int, float*, void*);
int unetca_tc_pack_pair(const void*, int, void*, int, int, void*);
int unetca_tc_conv3x3_fwd_rp64(const void*, int, const void*, int, void*, int, int, int, int, float*, void*);
int unetca_tc_conv3x3_fwd_split(const void*, int, const void*, int, void*, int, void*, int, int, int, int, int, int, int, float*, void*);
int unetca_tc_conv3x3_fwd_cat(const void*, int, const void*, int, int, const void*, void*, int, int, int, int, int, int, float*,
                              const float*, const float*, float*, void*);
int unetca_tc_conv3x3_wgrad_cat(const void*, int, const void*, int, const void*, int, int, float*, long, int, int, int, int, int, void*);
int unetca_tc_conv3x3_dgrad_bnstats(const void*, int, const void*, int, void*, int, int, int, int, int, int, const void*, int,
                                    const float*, const float*, const float*, float*, void*);
int unetca_tc_conv3x3_fwd_kw(const void*, int, const void*, void*, int, int, int, int, int, float*, void*);
int unetca_tc_pack_kw(const void*, int, void*, int, void*);
int unetca_tc_conv3x3_bnrelu_fwd(const void*, int, const void*, int, void*, int, int, int, int, int, int, const float*, const float*, float*, void*);
int unetca_tc_first_pairs_bnrelu_fwd(const void*, const void*, void*, int, int, int, int, int, const float*, const float*, void*);
int unetca_tc_first_pairs_fwd(const void*, const void*, void*, int, int, int, int, int, float*, void*);
int unetca_tc_first_pairs_wgrad(const void*, int, const void*, float*, long, int, int, int, int, void*);
int unetca_first_pairs_fold(const float*, int, int, int, float*, void*);
int unetca_tc_gemm_nt(const void*, int, const void*, int, void*, int, long, int, int, float*, void*);
int unetca_tc_convT_fwd(const void*, int, const void*, const float*, void*, int, int, int, int, int, int, void*);
int unetca_tc_convT_dgrad(const void*, int, const void*, void*, int, int, int, int, int, int, void*);
int unetca_tc_conv3x3_wgrad(const void*, int, const void*, int, float*, long, int, int, int, int, int, void*);
int unetca_tc_gemm_tn(const void*, int, const void*, int, float*, long, int, int, long, void*);
int unetca_tc_convT_wgrad(const void*, int, const void*, int, float*, long, int, int, int, int, int, void*);
int unetca_simt_conv3x3_fwd(int, const void*, int, const void*, int, void*, int, int, int, int, int, int, void*);
int unetca_simt_gemm_nt(int, const void*, int, const void*, int, void*, int, int, int, int, void*);
int unetca_simt_conv3x3_wgrad(int, const void*, int, const void*, int, float*, long, int, int, int, int, int, void*);
int unetca_simt_gemm_tn(int, const void*, int, const void*, int, float*, long, int, int, long, void*);
int unetca_simt_convT_fwd(int, const void*, int, const void*, const float*, void*, int, int, int, int, int, int, void*);
int unetca_simt_convT_dgrad(int, const void*, int, const void*, void*, int, int, int, int, int, int, void*);
int unetca_simt_convT_wgrad(int, const void*, int, const void*, int, float*, long, int, int, int, int, int, void*);
int unetca_chan_stats(int, const void*, int, int, long, float*, int*, void*);
int unetca_wgrad_reduce(const float*, int, long, int, int, int, int, float*, void*);
}

static int g_conv_impl = 0;   // 0: by dtype, 1: force FFMA kernels
static bool use_tc(int dtype) { return dtype == UNETCA_DTYPE_BF16 && g_conv_impl == 0; }

extern "C" {

void unetca_set_conv_impl(int impl) { g_conv_impl = impl; }
int unetca_get_conv_impl(void) { return g_conv_impl; }

// y = conv3x3(x, pad 1) without bias; NHWC, w packed [O][ldk] with k = tap*C + c.            UCA:81,84
// stat_parts (optional): partial per-channel sum / sum-of-squares of y, layout [*nparts][2][O].
// Also the dgrad: call with (dy, w_dgrad [C][9*O]) and C/O swapped.
int unetca_conv3x3_fwd(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W,
                       int C, int O, float* stat_parts, int* nparts, void* stream) {
    int rc;
    if (use_tc(dtype)) {
        rc = unetca_tc_conv3x3_fwd(x, ldx, w, ldk, y, ldy, B, H, W, C, O, stat_parts, stream);
        if (rc < 0) return rc;
        if (nparts) *nparts = rc;
        return 0;
    }
    rc = unetca_simt_conv3x3_fwd(dtype, x, ldx, w, ldk, y, ldy, B, H, W, C, O, stream);
    if (rc < 0) return rc;
    if (stat_parts) return unetca_chan_stats(dtype, y, ldy, O, (long)B * H * W, stat_parts, nparts, stream);
    return 0;
}

// The same convolution for narrow outputs (O a multiple of 64, H even) through the row-pair layout of the tcgen05
// path: w_pair [2*O][12*C] comes from unetca_pack_conv3x3_pair.  bf16 / tensor-core implementation only.
int unetca_conv3x3_fwd_paired(int dtype, const void* x, int ldx, const void* w_pair, void* y, int ldy, int B, int H, int W,
                              int C, int O, float* stat_parts, int* nparts, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_fwd_paired: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_conv3x3_fwd_paired(x, ldx, w_pair, y, ldy, B, H, W, C, O, stat_parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}

// 64 -> 64 channels, H even: resident-filter row-pair kernel on the ordinary packed filter [64][ldk] (bf16 tensor cores only)
int unetca_conv3x3_fwd_rp64(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W,
                            float* stat_parts, int* nparts, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_fwd_rp64: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_conv3x3_fwd_rp64(x, ldx, w, ldk, y, ldy, B, H, W, stat_parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}

// conv3x3 fwd / dgrad (O % 128 == 0) whose output channels go to TWO tensors: [0, split) -> y, [split, O) -> y2 (bf16 tensor cores)
int unetca_conv3x3_fwd_split(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, void* y2, int ldy2,
                             int split, int B, int H, int W, int C, int O, float* stat_parts, int* nparts, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_fwd_split: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_conv3x3_fwd_split(x, ldx, w, ldk, y, ldy, y2, ldy2, split, B, H, W, C, O, stat_parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}

// conv3x3 forward of torch.cat([x, x2], 1) (UCA:140) with the two halves as dense tensors; see the header.  bf16 tensor cores only.
int unetca_conv3x3_fwd_cat(int dtype, const void* x, int ldx, const void* x2, int ldx2, int C1, const void* w, void* y, int ldy,
                           int B, int H, int W, int C, int O, float* stat_parts, const float* scale, const float* shift,
                           float* sq_parts, int* nparts, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_fwd_cat: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_conv3x3_fwd_cat(x, ldx, x2, ldx2, C1, w, y, ldy, B, H, W, C, O, stat_parts, scale, shift, sq_parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}
int unetca_conv3x3_wgrad_cat(int dtype, const void* dy, int lddy, const void* x, int ldx, const void* x2, int ldx2, int C1,
                             float* ws, long ws_floats, int B, int H, int W, int C, int O, float* dw, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_wgrad_cat: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int ns = unetca_tc_conv3x3_wgrad_cat(dy, lddy, x, ldx, x2, ldx2, C1, ws, ws_floats, B, H, W, C, O, stream);
    if (ns < 0) return ns;
    return unetca_wgrad_reduce(ws, ns, (long)O * 9 * C, 0, O, C, 9 * C, dw, stream);
}

// dgrad + ReLU/BatchNorm backward statistics of its output in one kernel (replaces the unetca_bn_bwd_reduce pass)
int unetca_conv3x3_dgrad_bnstats(int dtype, const void* dy, int lddy, const void* wd, int ldk, void* da, int ldda, int B, int H,
                                 int W, int C, int O, const void* y1, int ldy1, const float* scale, const float* shift,
                                 const float* mean, float* parts, int* nparts, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_dgrad_bnstats: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_conv3x3_dgrad_bnstats(dy, lddy, wd, ldk, da, ldda, B, H, W, C, O, y1, ldy1, scale, shift, mean, parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}

// Inference: conv3x3 + folded eval-mode BatchNorm + ReLU in one tcgen05 kernel (bf16 tensor-core path only).
int unetca_conv3x3_bnrelu_fwd(int dtype, const void* x, int ldx, const void* w, int layout, void* y, int ldy, int B, int H,
                              int W, int C, int O, const float* scale, const float* shift, float* sq_parts, int* nparts,
                              void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_bnrelu_fwd: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_conv3x3_bnrelu_fwd(x, ldx, w, layout, y, ldy, B, H, W, C, O, scale, shift, sq_parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}
int unetca_first_pairs_bnrelu_fwd(int dtype, const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O,
                                  const float* scale, const float* shift, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("first_pairs_bnrelu_fwd: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    return unetca_tc_first_pairs_bnrelu_fwd(colp, wp, y, ldy, B, H, W, O, scale, shift, stream);
}

// 64 output channels, C = 64 / 128: kw-stacked layout (see conv_tc.cu).  bf16 / tensor-core implementation only.
int unetca_conv3x3_fwd_kw(int dtype, const void* x, int ldx, const void* w_kw, void* y, int ldy, int B, int H, int W, int C,
                          float* stat_parts, int* nparts, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("conv3x3_fwd_kw: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_conv3x3_fwd_kw(x, ldx, w_kw, y, ldy, B, H, W, C, stat_parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}
int unetca_pack_conv3x3_kw(int dtype, const void* w, int ld, void* w_kw, int C, void* stream) {
    if (dtype != UNETCA_DTYPE_BF16) { unetca::set_error("pack_conv3x3_kw: bf16 only"); return UNETCA_ERR_UNSUPPORTED; }
    return unetca_tc_pack_kw(w, ld, w_kw, C, stream);
}

// w [rows][ld] (K-major packed filter, k = tap*C + c, bf16) -> w_pair [2*rows][12*C]
int unetca_pack_conv3x3_pair(int dtype, const void* w, int ld, void* w_pair, int rows, int C, void* stream) {
    if (dtype != UNETCA_DTYPE_BF16) { unetca::set_error("pack_conv3x3_pair: bf16 only"); return UNETCA_ERR_UNSUPPORTED; }
    return unetca_tc_pack_pair(w, ld, w_pair, rows, C, stream);
}

// First conv through the pixel-pair layout (bf16 tensor-core path only; see unetca_im2col_pairs).
int unetca_first_pairs_fwd(int dtype, const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O,
                           float* stat_parts, int* nparts, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("first_pairs_fwd: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int rc = unetca_tc_first_pairs_fwd(colp, wp, y, ldy, B, H, W, O, stat_parts, stream);
    if (rc < 0) return rc;
    if (nparts) *nparts = rc;
    return 0;
}
int unetca_first_pairs_wgrad(int dtype, const void* dy, int lddy, const void* colp, float* ws, long ws_floats, int B, int H,
                             int W, int Cin, int O, float* dw, void* stream) {
    if (!use_tc(dtype)) { unetca::set_error("first_pairs_wgrad: bf16 tensor-core path only"); return UNETCA_ERR_UNSUPPORTED; }
    int ns = unetca_tc_first_pairs_wgrad(dy, lddy, colp, ws, ws_floats, B, H, W, O, stream);
    if (ns < 0) return ns;
    return unetca_first_pairs_fold(ws, ns, O, Cin, dw, stream);
}

// out[m][n] = sum_k A[m][k] * Bw[n][k]  (first conv on im2col rows, K = Kpad)
int unetca_gemm_nt(int dtype, const void* A, int lda, const void* Bw, int ldb, void* out, int ldo, long M, int N, int K,
                   float* stat_parts, int* nparts, void* stream) {
    int rc;
    if (use_tc(dtype)) {
        rc = unetca_tc_gemm_nt(A, lda, Bw, ldb, out, ldo, M, N, K, stat_parts, stream);
        if (rc < 0) return rc;
        if (nparts) *nparts = rc;
        return 0;
    }
    rc = unetca_simt_gemm_nt(dtype, A, lda, Bw, ldb, out, ldo, (int)M, N, K, stream);
    if (rc < 0) return rc;
    if (stat_parts) return unetca_chan_stats(dtype, out, ldo, N, M, stat_parts, nparts, stream);
    return 0;
}

// dW (O,C,3,3) fp32 = sum_p dy[p][o] * x[p+s(tap)][c]; ws = split-K scratch of ws_floats floats
int unetca_conv3x3_wgrad(int dtype, const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats, int B,
                         int H, int W, int C, int O, float* dw, void* stream) {
    int ns = use_tc(dtype) ? unetca_tc_conv3x3_wgrad(dy, lddy, x, ldx, ws, ws_floats, B, H, W, C, O, stream)
                           : unetca_simt_conv3x3_wgrad(dtype, dy, lddy, x, ldx, ws, ws_floats, B, H, W, C, O, stream);
    if (ns < 0) return ns;
    return unetca_wgrad_reduce(ws, ns, (long)O * 9 * C, 0, O, C, 9 * C, dw, stream);
}

// first conv: dW (O,Cin,3,3) fp32 = sum_p dy[p][o] * col[p][tap*Cin+c], col rows of width Kpad
int unetca_im2col_wgrad(int dtype, const void* dy, int lddy, const void* col, int Kpad, float* ws, long ws_floats,
                        long npix, int Cin, int O, float* dw, void* stream) {
    int ns = use_tc(dtype) ? unetca_tc_gemm_tn(dy, lddy, col, Kpad, ws, ws_floats, O, Kpad, npix, stream)
                           : unetca_simt_gemm_tn(dtype, dy, lddy, col, Kpad, ws, ws_floats, O, Kpad, npix, stream);
    if (ns < 0) return ns;
    return unetca_wgrad_reduce(ws, ns, (long)O * Kpad, 0, O, Cin, Kpad, dw, stream);
}

// ConvTranspose2d(k=2,s=2): out[b,2i+d,2j+e,o] = sum_c x[b,i,j,c] w[(d*2+e)*Cout+o][c] + bias[o]     UCA:112..121
// `out` may point into the upper channel half of a concat buffer (ldo = 2*Cout): torch.cat becomes free.
int unetca_convT2x2_fwd(int dtype, const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B,
                        int h, int wd, int Cin, int Cout, void* stream) {
    if (use_tc(dtype)) return unetca_tc_convT_fwd(x, ldx, w, bias, out, ldo, B, h, wd, Cin, Cout, stream);
    return unetca_simt_convT_fwd(dtype, x, ldx, w, bias, out, ldo, B, h, wd, Cin, Cout, stream);
}
int unetca_convT2x2_dgrad(int dtype, const void* dout, int ldd, const void* wdg, void* dx, int ldx, int B, int h, int wd,
                          int Cin, int Cout, void* stream) {
    if (use_tc(dtype)) return unetca_tc_convT_dgrad(dout, ldd, wdg, dx, ldx, B, h, wd, Cin, Cout, stream);
    return unetca_simt_convT_dgrad(dtype, dout, ldd, wdg, dx, ldx, B, h, wd, Cin, Cout, stream);
}
// dW (Cin,Cout,2,2) fp32
int unetca_convT2x2_wgrad(int dtype, const void* x, int ldx, const void* dout, int ldd, float* ws, long ws_floats, int B,
                          int h, int wd, int Cin, int Cout, float* dw, void* stream) {
    int ns = use_tc(dtype) ? unetca_tc_convT_wgrad(x, ldx, dout, ldd, ws, ws_floats, B, h, wd, Cin, Cout, stream)
                           : unetca_simt_convT_wgrad(dtype, x, ldx, dout, ldd, ws, ws_floats, B, h, wd, Cin, Cout, stream);
    if (ns < 0) return ns;
    return unetca_wgrad_reduce(ws, ns, (long)Cin * 4 * Cout, 1, Cin, Cout, 4 * Cout, dw, stream);
}

}  // extern "C"
