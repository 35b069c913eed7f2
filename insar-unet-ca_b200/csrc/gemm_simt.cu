// gemm_simt.cu — CUDA-core (FFMA, fp32 accumulate) implicit-GEMM contractions.
//
// This is the fp32 *parity mode* of the contractions on the U-Net-CA hot path (tcgen05 has no true-fp32 MMA,
// SURVEY.md §7.3), and the on-device cross-check of the tcgen05 kernels in conv_tc.cu.  Same operand layouts,
// same split-K workspace format, same entry-point signatures as the tensor-core path.
//
//   conv3x3 fwd / dgrad   UCA:81,84  (dgrad = the same contraction over the rotated, transposed filter)
//   conv3x3 wgrad         UCA:345
//   ConvTranspose2d k2s2 fwd / dgrad / wgrad   UCA:112,115,118,121
//   plain NT / TN GEMMs for the im2col'ed first conv (K = 9*Cin)
//
// One generic 64x64x16 tile kernel; a "problem" functor supplies A(m,k), B(n,k) and the epilogue store.
#include "common.cuh"

namespace unetca {

constexpr int TM = 64, TN = 64, TK = 16;

template <typename P>
__global__ void __launch_bounds__(256) simt_gemm_kernel(P prob, int M, int N, long K, long kchunk) {
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
    const long kbeg = (long)blockIdx.z * kchunk;
    long kend = kbeg + kchunk; if (kend > K) kend = K;
    const int tx = tid % 16, ty = tid / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long k0 = kbeg; k0 < kend; k0 += TK) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256;
            int ml, kl;
            if (P::A_KFAST) { kl = e % TK; ml = e / TK; } else { ml = e % TM; kl = e / TM; }
            const int m = m0 + ml; const long k = k0 + kl;
            As[kl][ml] = (m < M && k < kend) ? prob.a(m, k) : 0.f;
            int nl;
            if (P::B_KFAST) { kl = e % TK; nl = e / TK; } else { nl = e % TN; kl = e / TN; }
            const int n = n0 + nl; const long kb = k0 + kl;
            Bs[kl][nl] = (n < N && kb < kend) ? prob.b(n, kb) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) prob.store(m, n, acc[i][j], blockIdx.z);
        }
    }
}

// ---- problem functors ------------------------------------------------------------------------------------
template <typename T> struct Conv3x3Fwd {      // M = B*H*W, N = O, K = 9*C
    static constexpr bool A_KFAST = true, B_KFAST = true;
    const T* x; int ldx; const T* w; int ldk; T* y; int ldy; int H, W, C;
    __device__ float a(int m, long k) const {
        const int tap = (int)(k / C), c = (int)(k % C);
        const int wq = m % W, hq = (m / W) % H;
        const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) return 0.f;
        return to_float(x[((long)m + (long)(tap / 3 - 1) * W + (tap % 3 - 1)) * ldx + c]);
    }
    __device__ float b(int n, long k) const { return to_float(w[(long)n * ldk + k]); }
    __device__ void store(int m, int n, float v, int) const { y[(long)m * ldy + n] = from_float<T>(v); }
};
template <typename T> struct GemmNT {          // out[m][n] = sum_k A[m][k] B[n][k]
    static constexpr bool A_KFAST = true, B_KFAST = true;
    const T* A; int lda; const T* Bm; int ldb; T* out; int ldo;
    __device__ float a(int m, long k) const { return to_float(A[(long)m * lda + k]); }
    __device__ float b(int n, long k) const { return to_float(Bm[(long)n * ldb + k]); }
    __device__ void store(int m, int n, float v, int) const { out[(long)m * ldo + n] = from_float<T>(v); }
};
template <typename T> struct Conv3x3Wgrad {    // ws[z][o][tap*C+c] = sum_p dY[p][o] X[p+tap][c];  M=O, N=9C, K=npix
    static constexpr bool A_KFAST = false, B_KFAST = false;
    const T* dy; int lddy; const T* x; int ldx; float* ws; long split_stride; int ldn; int H, W, C;
    __device__ float a(int m, long k) const { return to_float(dy[k * lddy + m]); }
    __device__ float b(int n, long k) const {
        const int tap = n / C, c = n % C;
        const int wq = (int)(k % W), hq = (int)((k / W) % H);
        const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
        if (hh < 0 || hh >= H || ww < 0 || ww >= W) return 0.f;
        return to_float(x[(k + (long)(tap / 3 - 1) * W + (tap % 3 - 1)) * ldx + c]);
    }
    __device__ void store(int m, int n, float v, int z) const { ws[(long)z * split_stride + (long)m * ldn + n] = v; }
};
template <typename T> struct GemmTN {          // ws[z][m][n] = sum_k A[k][m] B[k][n]
    static constexpr bool A_KFAST = false, B_KFAST = false;
    const T* A; int lda; const T* Bm; int ldb; float* ws; long split_stride; int ldn;
    __device__ float a(int m, long k) const { return to_float(A[k * lda + m]); }
    __device__ float b(int n, long k) const { return to_float(Bm[k * ldb + n]); }
    __device__ void store(int m, int n, float v, int z) const { ws[(long)z * split_stride + (long)m * ldn + n] = v; }
};
template <typename T> struct ConvTFwd {        // M = B*h*w, N = 4*Cout, K = Cin; scatter to (2i+d, 2j+e)
    static constexpr bool A_KFAST = true, B_KFAST = true;
    const T* x; int ldx; const T* w; const float* bias; T* out; int ldo; int h, wd, Cin, Cout;
    __device__ float a(int m, long k) const { return to_float(x[(long)m * ldx + k]); }
    __device__ float b(int n, long k) const { return to_float(w[(long)n * Cin + k]); }
    __device__ void store(int m, int n, float v, int) const {
        const int de = n / Cout, o = n % Cout;
        const int j = m % wd, i = (m / wd) % h; const long bb = m / (wd * h);
        const long q = (bb * 2 * h + 2 * i + (de >> 1)) * 2 * wd + 2 * j + (de & 1);
        out[q * ldo + o] = from_float<T>(v + bias[o]);
    }
};
template <typename T> struct ConvTDgrad {      // M = B*h*w, N = Cin, K = 4*Cout
    static constexpr bool A_KFAST = true, B_KFAST = true;
    const T* dout; int ldd; const T* w; T* dx; int ldx; int h, wd, Cin, Cout;
    __device__ float a(int m, long k) const {
        const int de = (int)(k / Cout), o = (int)(k % Cout);
        const int j = m % wd, i = (m / wd) % h; const long bb = m / (wd * h);
        const long q = (bb * 2 * h + 2 * i + (de >> 1)) * 2 * wd + 2 * j + (de & 1);
        return to_float(dout[q * ldd + o]);
    }
    __device__ float b(int n, long k) const { return to_float(w[(long)n * 4 * Cout + k]); }
    __device__ void store(int m, int n, float v, int) const { dx[(long)m * ldx + n] = from_float<T>(v); }
};
template <typename T> struct ConvTWgrad {      // ws[z][cin][de*Cout+o] = sum_p X[p][cin] dOut[(p,de)][o]
    static constexpr bool A_KFAST = false, B_KFAST = false;
    const T* x; int ldx; const T* dout; int ldd; float* ws; long split_stride; int h, wd, Cin, Cout;
    __device__ float a(int m, long k) const { return to_float(x[k * ldx + m]); }
    __device__ float b(int n, long k) const {
        const int de = n / Cout, o = n % Cout;
        const int j = (int)(k % wd), i = (int)((k / wd) % h); const long bb = k / ((long)wd * h);
        const long q = (bb * 2 * h + 2 * i + (de >> 1)) * 2 * wd + 2 * j + (de & 1);
        return to_float(dout[q * ldd + o]);
    }
    __device__ void store(int m, int n, float v, int z) const {
        ws[(long)z * split_stride + (long)m * 4 * Cout + n] = v;
    }
};

template <typename P>
static int launch(const P& prob, int M, int N, long K, int nsplit, cudaStream_t st, const char* what) {
    long kchunk = (K + nsplit - 1) / nsplit;
    kchunk = (kchunk + TK - 1) / TK * TK;
    dim3 grid(ceil_div(M, TM), ceil_div(N, TN), ceil_div(K, kchunk));
    simt_gemm_kernel<P><<<grid, 256, 0, st>>>(prob, M, N, K, kchunk);
    int rc = check_launch(what);
    return rc < 0 ? rc : (int)grid.z;
}
// choose split-K so that the grid has ~2 waves of CTAs
static int pick_split(int M, int N, long K) {
    const long tiles = (long)ceil_div(M, TM) * ceil_div(N, TN);
    long s = (2L * num_sms() + tiles - 1) / tiles;
    const long maxs = (K + 4 * TK - 1) / (4 * TK);
    if (s > maxs) s = maxs;
    if (s < 1) s = 1;
    if (s > 512) s = 512;
    return (int)s;
}

}  // namespace unetca

using namespace unetca;
#define DISPATCH_T(dtype, ...)                                                          \
    do {                                                                                \
        if ((dtype) == UNETCA_DTYPE_F32) { typedef float T; __VA_ARGS__; }              \
        else if ((dtype) == UNETCA_DTYPE_BF16) { typedef bf16 T; __VA_ARGS__; }         \
        else { set_error("bad dtype %d", (int)(dtype)); return UNETCA_ERR_ARG; }        \
    } while (0)

extern "C" {

int unetca_simt_conv3x3_fwd(int dtype, const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H,
                            int W, int C, int O, void* stream) {
    DISPATCH_T(dtype, {
        Conv3x3Fwd<T> p{(const T*)x, ldx, (const T*)w, ldk, (T*)y, ldy, H, W, C};
        int rc = launch(p, B * H * W, O, 9L * C, 1, (cudaStream_t)stream, "simt_conv3x3_fwd");
        return rc < 0 ? rc : 0;
    });
    return UNETCA_OK;
}
int unetca_simt_gemm_nt(int dtype, const void* A, int lda, const void* Bm, int ldb, void* out, int ldo, int M, int N,
                        int K, void* stream) {
    DISPATCH_T(dtype, {
        GemmNT<T> p{(const T*)A, lda, (const T*)Bm, ldb, (T*)out, ldo};
        int rc = launch(p, M, N, K, 1, (cudaStream_t)stream, "simt_gemm_nt");
        return rc < 0 ? rc : 0;
    });
    return UNETCA_OK;
}
// returns the number of split-K partials written to ws ([nsplit][O][ldn], ldn = 9*C), or <0
int unetca_simt_conv3x3_wgrad(int dtype, const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats,
                              int B, int H, int W, int C, int O, void* stream) {
    DISPATCH_T(dtype, {
        const long K = (long)B * H * W;
        int ns = pick_split(O, 9 * C, K);
        const long stride = (long)O * 9 * C;
        if (ns * stride > ws_floats) ns = (int)(ws_floats / stride);
        if (ns < 1) { set_error("conv3x3_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
        Conv3x3Wgrad<T> p{(const T*)dy, lddy, (const T*)x, ldx, ws, stride, 9 * C, H, W, C};
        return launch(p, O, 9 * C, K, ns, (cudaStream_t)stream, "simt_conv3x3_wgrad");
    });
    return UNETCA_OK;
}
// ws[z][M][N] = sum_k A[k][m] B[k][n]; returns nsplit
int unetca_simt_gemm_tn(int dtype, const void* A, int lda, const void* Bm, int ldb, float* ws, long ws_floats, int M,
                        int N, long K, void* stream) {
    DISPATCH_T(dtype, {
        int ns = pick_split(M, N, K);
        const long stride = (long)M * N;
        if (ns * stride > ws_floats) ns = (int)(ws_floats / stride);
        if (ns < 1) { set_error("gemm_tn: workspace too small"); return UNETCA_ERR_WORKSPACE; }
        GemmTN<T> p{(const T*)A, lda, (const T*)Bm, ldb, ws, stride, N};
        return launch(p, M, N, K, ns, (cudaStream_t)stream, "simt_gemm_tn");
    });
    return UNETCA_OK;
}
int unetca_simt_convT_fwd(int dtype, const void* x, int ldx, const void* w, const float* bias, void* out, int ldo,
                          int B, int h, int wd, int Cin, int Cout, void* stream) {
    DISPATCH_T(dtype, {
        ConvTFwd<T> p{(const T*)x, ldx, (const T*)w, bias, (T*)out, ldo, h, wd, Cin, Cout};
        int rc = launch(p, B * h * wd, 4 * Cout, Cin, 1, (cudaStream_t)stream, "simt_convT_fwd");
        return rc < 0 ? rc : 0;
    });
    return UNETCA_OK;
}
int unetca_simt_convT_dgrad(int dtype, const void* dout, int ldd, const void* w, void* dx, int ldx, int B, int h,
                            int wd, int Cin, int Cout, void* stream) {
    DISPATCH_T(dtype, {
        ConvTDgrad<T> p{(const T*)dout, ldd, (const T*)w, (T*)dx, ldx, h, wd, Cin, Cout};
        int rc = launch(p, B * h * wd, Cin, 4L * Cout, 1, (cudaStream_t)stream, "simt_convT_dgrad");
        return rc < 0 ? rc : 0;
    });
    return UNETCA_OK;
}
int unetca_simt_convT_wgrad(int dtype, const void* x, int ldx, const void* dout, int ldd, float* ws, long ws_floats,
                            int B, int h, int wd, int Cin, int Cout, void* stream) {
    DISPATCH_T(dtype, {
        const long K = (long)B * h * wd;
        int ns = pick_split(Cin, 4 * Cout, K);
        const long stride = (long)Cin * 4 * Cout;
        if (ns * stride > ws_floats) ns = (int)(ws_floats / stride);
        if (ns < 1) { set_error("convT_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
        ConvTWgrad<T> p{(const T*)x, ldx, (const T*)dout, ldd, ws, stride, h, wd, Cin, Cout};
        return launch(p, Cin, 4 * Cout, K, ns, (cudaStream_t)stream, "simt_convT_wgrad");
    });
    return UNETCA_OK;
}

}  // extern "C"
