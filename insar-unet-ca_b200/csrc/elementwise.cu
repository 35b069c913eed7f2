// elementwise.cu — the HBM-bound half of the U-Net-CA hot path (everything that is not a contraction).
//
// Reference semantics restated here (file:line into /root/reference/Unet-ChannalAttention.py, "UCA"):
//   BatchNorm2d train/eval ......... UCA:82,85     ReLU ............... UCA:83,86
//   SELayer (pool, FC, sigmoid, scale) UCA:45-72   MaxPool2d(2) ....... UCA:106-109
//   outc 1x1 conv -> class logits ... UCA:125,162  CrossEntropyLoss(ignore_index=255) UCA:465,344
//   torch.max(outputs,1) argmax mask  UCA:220
// plus the autograd derivatives of each (UCA:345).
//
// All kernels share one thread mapping ("pixel rows"): a 256-thread block owns a contiguous range of
// NHWC pixels; thread (r, cv) walks pixels r, r+rows, ... and always touches the same 16-byte channel
// vector cv, so per-channel parameters are loaded once per thread and every warp-level access is a run of
// whole 128-byte pixel rows (fully coalesced).  Reductions are two-stage and deterministic: per-block
// partials in a caller-provided fp32 scratch, then a tiny finalize kernel that sums them in double.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <math.h>

namespace unetca {
int make_tmap_nhwc(CUtensorMap* m, const void* base, int elem_bytes, long C, long W, long H, long B, long ld, int bw, int bh);
}

namespace unetca {

constexpr int kThreads = 256;
constexpr int kMaxParts = 1184;  // 148 SMs x 8

// tuning knobs (unetca_set_tuning): pixels per thread-row of an elementwise block; waves of a per-image reduction grid
// (0 = legacy sizing by kMaxParts)
static int g_ew_px = 16;
static int g_apply_stream = 8;      // > 0: BN-backward / squeeze passes as shared-memory streams; bn_bwd_apply: this many 8 KB tiles per block
static int g_red_waves = 1;
static int g_pool_quads = 4;      // 2x2 quads per thread-row of a se_scale_pool block

struct RowMap {
    int vpr;    // 16-byte vectors per pixel row
    int rows;   // pixel rows processed concurrently by one block
};
template <typename T> static inline RowMap row_map(int C) {
    RowMap m;
    m.vpr = C / VecTraits<T>::N;
    m.rows = kThreads / m.vpr;
    return m;
}
// pixels per block for an elementwise pass / a reduction pass
template <typename T> static inline long ew_chunk(int C, long npix) {
    (void)npix;
    return (long)row_map<T>(C).rows * g_ew_px;
}
template <typename T> static inline long red_chunk(int C, long npix) {
    long c = (long)row_map<T>(C).rows * 16;
    long c2 = (npix + kMaxParts - 1) / kMaxParts;
    return c > c2 ? c : c2;
}

// Per-image reduction grids (grid = (nblk, B)): size the grid to the kernel's resident capacity (148 SMs x blocks/SM
// x g_red_waves) so that every block runs concurrently and all finish together; a grid of ~2 waves of large
// blocks loses up to a third of the bandwidth in its tail.
template <typename K> static int resident_blocks(K kernel) {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kThreads, 0) != cudaSuccess || v < 1) v = 1;
    return v * num_sms();
}
template <typename T> static inline long img_red_chunk(int C, long pix_per_img, int B, int slots) {
    if (g_red_waves <= 0) return red_chunk<T>(C, pix_per_img * B);
    const long rows = row_map<T>(C).rows;
    long cap = (long)slots * g_red_waves;
    if (cap > kMaxParts) cap = kMaxParts;
    long per_img = cap / B;
    if (per_img < 1) per_img = 1;
    long chunk = (pix_per_img + per_img - 1) / per_img;
    chunk = (chunk + rows - 1) / rows * rows;
    if (chunk < rows * 16) chunk = rows * 16;
    return chunk;
}

// Sum `acc[s][*]` over the block's pixel rows and write out[s*C + c]; all threads must call.
template <int NSTAT, int VEC>
__device__ __forceinline__ void block_reduce_rows(float (&acc)[NSTAT][VEC], int C, int vpr, int rows, float* out) {
    __shared__ float red[kThreads * 8];
    const int tid = threadIdx.x;
    const int r = tid / vpr, cv = tid % vpr;
#pragma unroll
    for (int s = 0; s < NSTAT; ++s) {
        __syncthreads();
        if (r < rows) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) red[r * C + cv * VEC + i] = acc[s][i];
        }
        __syncthreads();
        for (int c = tid; c < C; c += kThreads) {
            float t = 0.f;
            for (int rr = 0; rr < rows; ++rr) t += red[rr * C + c];
            out[s * C + c] = t;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// per-channel sum / sum of squares of a conv output (used when the producing GEMM does not fuse them)
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) chan_stats_kernel(const T* __restrict__ y, int ld, int C, long npix,
                                                              long chunk, float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > npix) p1 = npix;
    float acc[2][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
    if (r < rows) {
#pragma unroll 4
        for (long p = p0 + r; p < p1; p += rows) {
            float v[VEC];
            load_vec(y + p * ld + cv * VEC, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) { acc[0][i] += v[i]; acc[1][i] += v[i] * v[i]; }
        }
    }
    block_reduce_rows<2, VEC>(acc, C, vpr, rows, parts + (long)blockIdx.x * 2 * C);
}


// Sum NSTAT interleaved partial rows parts[i][s][C] over i for 32 consecutive channels per block: thread (cl, pl) walks
// parts pl, pl+8, ... (coalesced 128-byte reads), then the 8 part-lanes are folded through shared memory in double.
// blockDim = 256, grid = ceil(C/32).  Returns the totals of channel c = blockIdx.x*32 + (tid & 31) to threads < 32.
template <int NSTAT>
__device__ __forceinline__ bool sum_parts32(const float* __restrict__ parts, int nparts, int C, double (&tot)[NSTAT]) {
    __shared__ double red[NSTAT][8][32];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double acc[NSTAT];
#pragma unroll
    for (int s = 0; s < NSTAT; ++s) acc[s] = 0.0;
    if (c < C) {
        for (int i = pl; i < nparts; i += 8) {
#pragma unroll
            for (int s = 0; s < NSTAT; ++s) acc[s] += (double)parts[((long)i * NSTAT + s) * C + c];
        }
    }
#pragma unroll
    for (int s = 0; s < NSTAT; ++s) red[s][pl][cl] = acc[s];
    __syncthreads();
    if (pl != 0 || c >= C) return false;
#pragma unroll
    for (int s = 0; s < NSTAT; ++s) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[s][k][cl];
        tot[s] = t;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------------------
// BN finalize (train): batch mean / biased var from partial sums, running-stat update with the unbiased
// variance (momentum 0.1), and the fused affine a = gamma*invstd, b = beta - mean*a.      UCA:82,85
// The pre-BN conv bias never enters the conv: in train mode BN cancels it, so it only shifts the batch mean
// that goes into running_mean.
// ---------------------------------------------------------------------------------------------------------
__global__ void bn_finalize_train_kernel(const float* __restrict__ parts, int nparts, int C, double count,
                                         const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float* running_mean, float* running_var,
                                         float momentum, float eps, float* mean_out, float* invstd_out,
                                         float* scale_out, float* shift_out) {
    double tot[2];
    if (!sum_parts32<2>(parts, nparts, C, tot)) return;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const double s = tot[0], q = tot[1];
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = gamma[c] * invstd;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    scale_out[c] = a;
    shift_out[c] = beta[c] - (float)mean * a;
    if (running_mean) {
        const float bm = (float)mean + (conv_bias ? conv_bias[c] : 0.f);
        const float uv = (float)(var * (count / (count - 1.0)));
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * bm;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * uv;
    }
}

// eval: a = gamma / sqrt(running_var + eps), b = beta + (conv_bias - running_mean) * a    UCA:276
__global__ void bn_fold_eval_kernel(int C, const float* __restrict__ conv_bias, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ running_mean,
                                    const float* __restrict__ running_var, float eps, float* scale_out,
                                    float* shift_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float a = gamma[c] / sqrtf(running_var[c] + eps);
    scale_out[c] = a;
    shift_out[c] = beta[c] + ((conv_bias ? conv_bias[c] : 0.f) - running_mean[c]) * a;
}

// ---------------------------------------------------------------------------------------------------------
// out = relu(a*y + b)                      (first conv of a DoubleConv, or second conv when use_se=False)
// POOLSUM: additionally / instead accumulate per-(image, channel) sums of relu(a*y+b) for the SE squeeze.
//   grid = (blocks per image, B); WRITE selects whether `out` is written.
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool WRITE, bool POOLSUM>
__global__ void __launch_bounds__(kThreads) bn_relu_kernel(const T* __restrict__ y, int ldy, T* __restrict__ out,
                                                           int ldo, int C, long pix_per_img, long chunk,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long base = (long)blockIdx.y * pix_per_img;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    float acc[1][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = 0.f;
    if (r < rows) {
        float a[VEC], b[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { a[i] = scale[cv * VEC + i]; b[i] = shift[cv * VEC + i]; }
#pragma unroll 4
        for (long p = p0 + r; p < p1; p += rows) {
            float v[VEC];
            load_vec(y + (base + p) * ldy + cv * VEC, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                v[i] = fmaxf(fmaf(a[i], v[i], b[i]), 0.f);
                if (POOLSUM) acc[0][i] += v[i];
            }
            if (WRITE) store_vec(out + (base + p) * ldo + cv * VEC, v);
        }
    }
    if (POOLSUM)
        block_reduce_rows<1, VEC>(acc, C, vpr, rows, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * C);
}

// ---------------------------------------------------------------------------------------------------------
// Per-image partial rows -> one row per image, in place (row 0 of each image).  The per-image reduction grids write
// ~resident-capacity / B rows per image: 6 at batch 64, but 440 for a single large image, which one FC block per image
// would then sum alone (390 us at C = 1024).  Block = 32 columns x 8 row groups, grid = (ceil(width / 32), B); fixed
// summation order (row groups, then rows), double accumulation.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) img_parts_sum_kernel(float* __restrict__ parts, int nparts, int width) {
    __shared__ double red[8][32];
    const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + cl;
    float* base = parts + (long)blockIdx.y * nparts * width;
    double acc = 0.0;
    if (col < width) {
        int i = g;
        for (; i + 24 < nparts; i += 32) {
            const float a0 = base[(long)i * width + col], a1 = base[(long)(i + 8) * width + col];
            const float a2 = base[(long)(i + 16) * width + col], a3 = base[(long)(i + 24) * width + col];
            acc += (double)a0; acc += (double)a1; acc += (double)a2; acc += (double)a3;
        }
        for (; i < nparts; i += 8) acc += (double)base[(long)i * width + col];
    }
    red[g][cl] = acc;
    __syncthreads();
    if (g == 0 && col < width) {
        double t = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k][cl];
        base[col] = (float)t;
    }
}
// rows per image beyond which the FC kernels get the pre-reduced single row (launch + ~5 us against nparts sequential loads)
constexpr int kPartsPreReduce = 16;

// ---------------------------------------------------------------------------------------------------------
// SE excitation: p = mean_hw, z = relu(W1 p), s = sigmoid(W2 z)      UCA:54-59,65-68   (one block per image)
// parts rows of image b start at row b * img_rows; the first nparts of them are summed.
// ---------------------------------------------------------------------------------------------------------
template <int NSTAT>
__global__ void __launch_bounds__(256) se_fc_kernel(const float* __restrict__ parts, int nparts, int img_rows, int C, int Cr,
                                                    float inv_hw, const float* __restrict__ w1,
                                                    const float* __restrict__ w2, float* __restrict__ p_out,
                                                    float* __restrict__ z_out, float* __restrict__ s_out,
                                                    float* __restrict__ sums34, const float* __restrict__ scale,
                                                    const float* __restrict__ shift, const float* __restrict__ mean) {
    extern __shared__ float sm[];
    float* p = sm;          // [C]
    float* z = sm + C;      // [Cr]
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x) {
        double t = 0.0, t3 = 0.0, ty = 0.0;
        for (int i = 0; i < nparts; ++i) {
            if (NSTAT == 1) t += (double)parts[((long)b * img_rows + i) * C + c];
            else {
                const float* row = parts + ((long)b * img_rows + i) * 2 * C + c;
                t3 += (double)row[0];
                ty += (double)row[C];
            }
        }
        if (NSTAT == 3) t = (double)scale[c] * ty + (double)shift[c] * t3;       // sum relu(a*y+b) = a*Sy + b*S3
        const float v = (float)(t * (double)inv_hw);
        p[c] = v;
        p_out[(long)b * C + c] = v;
        if (NSTAT == 3 && sums34) {
            sums34[((long)b * 2 + 0) * C + c] = (float)t3;
            sums34[((long)b * 2 + 1) * C + c] = (float)(ty - (mean ? (double)mean[c] : 0.0) * t3);
        }
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < Cr; j += nwarps) {
        float t = 0.f;
        for (int c = lane; c < C; c += 32) t = fmaf(w1[(long)j * C + c], p[c], t);
        t = warp_sum(t);
        if (lane == 0) { t = fmaxf(t, 0.f); z[j] = t; z_out[(long)b * Cr + j] = t; }
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        float t = 0.f;
        if ((Cr & 3) == 0) {
            // the thread's own row of W2 as 16-byte loads (Cr = C / 16 is a multiple of 4 for every layer of the model); same
            // summation order as the scalar loop
            const float4* wr = reinterpret_cast<const float4*>(w2 + (long)c * Cr);
            for (int j4 = 0; j4 < Cr / 4; ++j4) {
                const float4 w = __ldg(wr + j4);
                t = fmaf(w.x, z[4 * j4], t); t = fmaf(w.y, z[4 * j4 + 1], t);
                t = fmaf(w.z, z[4 * j4 + 2], t); t = fmaf(w.w, z[4 * j4 + 3], t);
            }
        } else {
            for (int j = 0; j < Cr; ++j) t = fmaf(w2[(long)c * Cr + j], z[j], t);
        }
        s_out[(long)b * C + c] = 1.f / (1.f + expf(-t));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Block output: o = relu(a*y+b) * s[b,c]  (s == nullptr: plain relu(bn)), written with stride ldo (e.g. straight
// into the skip half of the decoder's concat buffer -> torch.cat at UCA:140.. costs nothing), and optionally
// the fused MaxPool2d(2) of o: pooled value + 1-byte window position (dh*2+dw).       UCA:72, UCA:106-109
// Tie-break / NaN rule of nn.MaxPool2d: scan (0,0),(0,1),(1,0),(1,1); take when v > best || isnan(v).
// One thread = one 2x2 pixel quad x one channel vector.
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool POOL>
__global__ void __launch_bounds__(kThreads) se_scale_pool_kernel(const T* __restrict__ y, int ldy,
                                                                 T* __restrict__ out, int ldo,
                                                                 T* __restrict__ pooled, int ldp,
                                                                 uint8_t* __restrict__ pos, int H, int W, int C,
                                                                 long chunk, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 const float* __restrict__ s) {
    // grid = (blocks per image, B); thread (r, cv) walks quads q0+r, q0+r+rows, ... of its image with one channel vector
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    if (r >= rows) return;
    const int Ho = H >> 1, Wo = W >> 1;
    const long nquad = (long)Ho * Wo;
    const int b = blockIdx.y;
    const long q0 = (long)blockIdx.x * chunk;
    long q1 = q0 + chunk; if (q1 > nquad) q1 = nquad;
    float a[VEC], sh[VEC], g[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        a[i] = scale[cv * VEC + i];
        sh[i] = shift[cv * VEC + i];
        g[i] = s ? s[(long)b * C + cv * VEC + i] : 1.f;
    }
#pragma unroll 2
    for (long q = q0 + r; q < q1; q += rows) {
        const int wo = (int)(q % Wo), ho = (int)(q / Wo);
        const long pbase = ((long)b * H + 2 * ho) * W + 2 * wo;
        float v[4][VEC];
#pragma unroll
        for (int k = 0; k < 4; ++k) load_vec(y + (pbase + (k >> 1) * W + (k & 1)) * ldy + cv * VEC, v[k]);
        float best[VEC];
        uint8_t code[VEC];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float t = fmaxf(fmaf(a[i], v[k][i], sh[i]), 0.f) * g[i];
                t = round_to(t, (const T*)nullptr);
                v[k][i] = t;
                if (POOL) {
                    if (k == 0) { best[i] = t; code[i] = 0; }
                    else if (t > best[i] || t != t) { best[i] = t; code[i] = (uint8_t)k; }
                }
            }
            store_vec(out + (pbase + (k >> 1) * W + (k & 1)) * ldo + cv * VEC, v[k]);
        }
        if (POOL) {
            const long qg = (long)b * nquad + q;
            store_vec(pooled + qg * ldp + cv * VEC, best);
            uint8_t* dst = pos + qg * C + cv * VEC;
            if (VEC == 8) {
                uint2 t;
                t.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
                t.y = code[4 % VEC] | (code[5 % VEC] << 8) | (code[6 % VEC] << 16) | (code[7 % VEC] << 24);
                *reinterpret_cast<uint2*>(dst) = t;
            } else {
                *reinterpret_cast<uint32_t*>(dst) = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
            }
        }
    }
}

// out = relu(a*y+b) * s[b,c] without pooling (any H, W): plain pixel-row pass; grid = (blocks per image, B)
template <typename T>
__global__ void __launch_bounds__(kThreads) se_scale_kernel(const T* __restrict__ y, int ldy, T* __restrict__ out, int ldo,
                                                            int C, long pix_per_img, long chunk,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift,
                                                            const float* __restrict__ s) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    if (r >= rows) return;
    const long base = (long)blockIdx.y * pix_per_img;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    float a[VEC], b[VEC], g[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        a[i] = scale[cv * VEC + i]; b[i] = shift[cv * VEC + i];
        g[i] = s ? s[(long)blockIdx.y * C + cv * VEC + i] : 1.f;
    }
#pragma unroll 4
    for (long p = p0 + r; p < p1; p += rows) {
        float v[VEC];
        load_vec(y + (base + p) * ldy + cv * VEC, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = fmaxf(fmaf(a[i], v[i], b[i]), 0.f) * g[i];
        store_vec(out + (base + p) * ldo + cv * VEC, v);
    }
}

// Standalone MaxPool2d(2) over an NHWC tensor: pooled values, 1-byte window positions and (optionally) the
// int64 flat indices h*W+w that torch's max_pool2d(return_indices=True) reports.          UCA:106-109
template <typename T>
__global__ void __launch_bounds__(kThreads) maxpool_kernel(const T* __restrict__ x, int ldx, T* __restrict__ pooled,
                                                           int ldp, uint8_t* __restrict__ pos,
                                                           long long* __restrict__ idx64, int B, int H, int W, int C) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC;
    const int Ho = H >> 1, Wo = W >> 1;
    const long nquad = (long)B * Ho * Wo;
    const long gid = (long)blockIdx.x * kThreads + threadIdx.x;
    const long q = gid / vpr;
    const int cv = (int)(gid % vpr);
    if (q >= nquad) return;
    const int wo = (int)(q % Wo);
    const int ho = (int)((q / Wo) % Ho);
    const int b = (int)(q / ((long)Wo * Ho));
    float best[VEC];
    int code[VEC];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long p = ((long)b * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1);
        float v[VEC];
        load_vec(x + p * ldx + cv * VEC, v);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            if (k == 0) { best[i] = v[i]; code[i] = 0; }
            else if (v[i] > best[i] || v[i] != v[i]) { best[i] = v[i]; code[i] = k; }
        }
    }
    store_vec(pooled + q * ldp + cv * VEC, best);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        if (pos) pos[q * C + cv * VEC + i] = (uint8_t)code[i];
        // NCHW-shaped (B,C,Ho,Wo) int64 output, value = flat offset into the (H,W) input plane
        if (idx64)
            idx64[(((long)b * C + cv * VEC + i) * Ho + ho) * Wo + wo] =
                (long long)(2 * ho + (code[i] >> 1)) * W + 2 * wo + (code[i] & 1);
    }
}

// dx = skip_grad + unpool(dpooled): the gradient of a skip tensor that feeds both the decoder concat and the
// next level's MaxPool.  skip_grad may be null (0).  One thread = one 2x2 quad x one channel vector.
template <typename T>
__global__ void __launch_bounds__(kThreads) pool_bwd_add_kernel(const T* __restrict__ skip_grad, int lds,
                                                                const T* __restrict__ dpooled, int ldp,
                                                                const uint8_t* __restrict__ pos, T* __restrict__ dx,
                                                                int ldx, int B, int H, int W, int C) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC;
    const int Ho = H >> 1, Wo = W >> 1;
    const long nquad = (long)B * Ho * Wo;
    const long gid = (long)blockIdx.x * kThreads + threadIdx.x;
    const long q = gid / vpr;
    const int cv = (int)(gid % vpr);
    if (q >= nquad) return;
    const int wo = (int)(q % Wo);
    const int ho = (int)((q / Wo) % Ho);
    const int b = (int)(q / ((long)Wo * Ho));
    float g[VEC];
    load_vec(dpooled + q * ldp + cv * VEC, g);
    uint8_t code[VEC];
    const uint8_t* src = pos + q * C + cv * VEC;
    if (VEC == 8) {
        uint2 t = *reinterpret_cast<const uint2*>(src);
#pragma unroll
        for (int i = 0; i < 4; ++i) { code[i] = (t.x >> (8 * i)) & 0xff; code[(4 + i) % VEC] = (t.y >> (8 * i)) & 0xff; }
    } else {
        uint32_t t = *reinterpret_cast<const uint32_t*>(src);
#pragma unroll
        for (int i = 0; i < 4; ++i) code[i] = (t >> (8 * i)) & 0xff;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const long p = ((long)b * H + 2 * ho + (k >> 1)) * W + 2 * wo + (k & 1);
        float v[VEC];
        if (skip_grad) load_vec(skip_grad + p * lds + cv * VEC, v);
        else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) v[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) if (code[i] == k) v[i] += g[i];
        store_vec(dx + p * ldx + cv * VEC, v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// SE backward, stage 1: ds[b,c] = sum_hw dO * relu(a*y+b)         (partials per image; grid = (nblk, B))
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) se_bwd_reduce_kernel(const T* __restrict__ dout, int ldd,
                                                                 const T* __restrict__ y, int ldy, int C,
                                                                 long pix_per_img, long chunk,
                                                                 const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long base = (long)blockIdx.y * pix_per_img;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    float acc[1][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = 0.f;
    if (r < rows) {
        float a[VEC], b[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { a[i] = scale[cv * VEC + i]; b[i] = shift[cv * VEC + i]; }
#pragma unroll 4
        for (long p = p0 + r; p < p1; p += rows) {
            float v[VEC], d[VEC];
            load_vec(y + (base + p) * ldy + cv * VEC, v);
            load_vec(dout + (base + p) * ldd + cv * VEC, d);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[0][i] = fmaf(d[i], fmaxf(fmaf(a[i], v[i], b[i]), 0.f), acc[0][i]);
        }
    }
    block_reduce_rows<1, VEC>(acc, C, vpr, rows, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * C);
}

// SE backward, stage 2 (one block per image): ds -> d(pre-sigmoid) -> dz -> dp.    derivative of UCA:54-59
//   dpre2[b,c] = ds*s*(1-s);  dz[b,j] = (sum_c dpre2[c] W2[c,j]) * (z>0);  dp[b,c] = sum_j dz[j] W1[j,c]
__global__ void __launch_bounds__(256) se_fc_bwd_kernel(const float* __restrict__ parts, int nparts, int C, int Cr,
                                                        const float* __restrict__ w1, const float* __restrict__ w2,
                                                        const float* __restrict__ z, const float* __restrict__ s,
                                                        float* __restrict__ dpre2_out, float* __restrict__ dz_out,
                                                        float* __restrict__ dp_out) {
    extern __shared__ float sm[];
    float* dpre2 = sm;       // [C]
    float* dz = sm + C;      // [Cr]
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x) {
        double t = 0.0;
        for (int i = 0; i < nparts; ++i) t += (double)parts[((long)b * nparts + i) * C + c];
        const float sv = s[(long)b * C + c];
        const float v = (float)t * sv * (1.f - sv);
        dpre2[c] = v;
        dpre2_out[(long)b * C + c] = v;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int j = warp; j < Cr; j += nwarps) {
        float t = 0.f;
        for (int c = lane; c < C; c += 32) t = fmaf(dpre2[c], w2[(long)c * Cr + j], t);
        t = warp_sum(t);
        if (lane == 0) {
            t = z[(long)b * Cr + j] > 0.f ? t : 0.f;
            dz[j] = t;
            dz_out[(long)b * Cr + j] = t;
        }
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int j = 0; j < Cr; ++j) t = fmaf(dz[j], w1[(long)j * C + c], t);
        dp_out[(long)b * C + c] = t;
    }
}

// SE FC weight grads: dW2[c,j] = sum_b dpre2[b,c] z[b,j];  dW1[j,c] = sum_b dz[b,j] p[b,c]   (K = batch)
__global__ void se_fc_wgrad_kernel(int B, int C, int Cr, const float* __restrict__ dpre2,
                                   const float* __restrict__ dz, const float* __restrict__ p,
                                   const float* __restrict__ z, float* __restrict__ dw1, float* __restrict__ dw2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= C * Cr) return;
    {   // dW2 is (C, Cr)
        const int c = i / Cr, j = i % Cr;
        float t = 0.f;
        for (int b = 0; b < B; ++b) t = fmaf(dpre2[(long)b * C + c], z[(long)b * Cr + j], t);
        dw2[i] = t;
    }
    {   // dW1 is (Cr, C)
        const int j = i / C, c = i % C;
        float t = 0.f;
        for (int b = 0; b < B; ++b) t = fmaf(dz[(long)b * Cr + j], p[(long)b * C + c], t);
        dw1[i] = t;
    }
}

// ---------------------------------------------------------------------------------------------------------
// ReLU + BN backward (train).  dz = dA * (a*y+b > 0) with, when SE follows (s != null),
// dA = dO*s[b,c] + dp[b,c]/HW (the SE scale and squeeze derivatives folded in).
//   stage 1: per-channel sum(dz), sum(dz*xhat)      stage 2 (finalize): dgamma, dbeta, c1, c2
//   stage 3: dY = gamma*invstd * (dz - c1 - xhat*c2)
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool SE, bool APPLY>
__global__ void __launch_bounds__(kThreads, APPLY ? 3 : 4) bn_bwd_kernel(const T* __restrict__ dout, int ldd, const T* __restrict__ y,
                                                          int ldy, T* __restrict__ dy, int lddy, int C,
                                                          long pix_per_img, long chunk, float inv_hw,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ invstd,
                                                          const float* __restrict__ s, const float* __restrict__ dp,
                                                          const float* __restrict__ coef /* [3][C]: g, c1, c2 */,
                                                          float* __restrict__ parts) {
    // reduce:  acc0 = sum dz, acc1 = sum dz*(y-mean)            (invstd is applied by the finalize kernel)
    // apply:   dY = g*dz - y*k2 + k0  with  k2 = g*c2*invstd,  k0 = mean*k2 - g*c1
    //          and, when SE follows, g*dz = mask * (dO*(g*s) + g*dp/HW)
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long base = (long)blockIdx.y * pix_per_img;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    float acc[2][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
    if (r < rows) {
        float a[VEC], b[VEC], m0[VEC], m1[VEC], k2[VEC], k0[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int c = cv * VEC + i;
            a[i] = scale[c]; b[i] = shift[c];
            float sv = 1.f, dpv = 0.f;
            if (SE) { sv = s[(long)blockIdx.y * C + c]; dpv = dp[(long)blockIdx.y * C + c] * inv_hw; }
            if (APPLY) {
                const float g = coef[c], c1 = coef[C + c], c2 = coef[2 * C + c];
                k2[i] = g * c2 * invstd[c];
                k0[i] = mean[c] * k2[i] - g * c1;
                m0[i] = g * sv; m1[i] = g * dpv;
            } else {
                m0[i] = sv; m1[i] = dpv; k2[i] = mean[c]; k0[i] = 0.f;
            }
        }
#pragma unroll 4
        for (long p = p0 + r; p < p1; p += rows) {
            float v[VEC], d[VEC];
            load_vec(y + (base + p) * ldy + cv * VEC, v);
            load_vec(dout + (base + p) * ldd + cv * VEC, d);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float da = SE ? fmaf(d[i], m0[i], m1[i]) : (APPLY ? d[i] * m0[i] : d[i]);
                const float dz = fmaf(a[i], v[i], b[i]) > 0.f ? da : 0.f;
                if (APPLY) d[i] = fmaf(-v[i], k2[i], dz) + k0[i];
                else { acc[0][i] += dz; acc[1][i] = fmaf(dz, v[i] - k2[i], acc[1][i]); }
            }
            if (APPLY) store_vec(dy + (base + p) * lddy + cv * VEC, d);
        }
    }
    if (!APPLY)
        block_reduce_rows<2, VEC>(acc, C, vpr, rows, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// ---------------------------------------------------------------------------------------------------------
// The APPLY pass as a shared-memory stream (dense bf16 tensors, C/8 a power of two <= 256): one elected thread moves
// 8 KB tiles of Y and dO with cp.async.bulk into a ring of stages, all threads transform a tile in place in shared
// memory, and the result leaves by a bulk store.  The register-based kernel above keeps only 2-4 16-byte loads per
// thread in flight (48 per-channel constants crowd the register file at 3 blocks/SM): ~25-50 KB per SM, at the edge of
// what HBM latency x bandwidth needs (~35 KB).  Here (kStStages - 1) x 16 KB per block are in flight whatever the
// register budget.  Same arithmetic in the same order: bit-identical results.
// ---------------------------------------------------------------------------------------------------------
constexpr int kStTile = 8192;        // bytes per operand per stage
constexpr int kStStages = 4;
constexpr int kStSmemBytes = kStStages * 2 * kStTile;

__device__ __forceinline__ void bulk_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* dst, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}

template <bool SE>
__global__ void __launch_bounds__(kThreads, 3) bn_bwd_apply_stream_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y,
                                                                          bf16* __restrict__ dy, int C, long pix_per_img, long chunk,
                                                                          float inv_hw, const float* __restrict__ scale,
                                                                          const float* __restrict__ shift,
                                                                          const float* __restrict__ mean,
                                                                          const float* __restrict__ invstd,
                                                                          const float* __restrict__ s, const float* __restrict__ dp,
                                                                          const float* __restrict__ coef) {
    extern __shared__ __align__(128) uint8_t st_smem[];        // [stage][y tile | dO tile]
    __shared__ __align__(8) uint64_t full_bar[kStStages];
    constexpr int VEC = 8;
    const int vpr = C / VEC;
    const int cv = threadIdx.x & (vpr - 1);
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    const long off0 = ((long)blockIdx.y * pix_per_img + p0) * C * 2;        // byte offset of this block's range
    const long nbytes = (p1 - p0) * C * 2;
    const int ntiles = (int)((nbytes + kStTile - 1) / kStTile);
    const uint8_t* yb = reinterpret_cast<const uint8_t*>(y) + off0;
    const uint8_t* db = reinterpret_cast<const uint8_t*>(dout) + off0;
    uint8_t* ob = reinterpret_cast<uint8_t*>(dy) + off0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int stg = t % kStStages;
        const long o = (long)t * kStTile;
        const uint32_t bytes = (uint32_t)(nbytes - o < kStTile ? nbytes - o : kStTile);
        mbar_expect_tx(&full_bar[stg], 2 * bytes);
        bulk_load_1d(st_smem + stg * 2 * kStTile, yb + o, bytes, &full_bar[stg]);
        bulk_load_1d(st_smem + stg * 2 * kStTile + kStTile, db + o, bytes, &full_bar[stg]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kStStages - 1 && t < ntiles; ++t) issue(t);
    float a[VEC], b[VEC], m0[VEC], m1[VEC], k2[VEC], k0[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = cv * VEC + i;
        a[i] = scale[c]; b[i] = shift[c];
        float sv = 1.f, dpv = 0.f;
        if (SE) { sv = s[(long)blockIdx.y * C + c]; dpv = dp[(long)blockIdx.y * C + c] * inv_hw; }
        const float g = coef[c], c1 = coef[C + c], c2 = coef[2 * C + c];
        k2[i] = g * c2 * invstd[c];
        k0[i] = mean[c] * k2[i] - g * c1;
        m0[i] = g * sv; m1[i] = g * dpv;
    }
    for (int t = 0; t < ntiles; ++t) {
        const int stg = t % kStStages;
        const long o = (long)t * kStTile;
        const int bytes = (int)(nbytes - o < kStTile ? nbytes - o : kStTile);
        mbar_wait(&full_bar[stg], (uint32_t)((t / kStStages) & 1));
        bf16* ys = reinterpret_cast<bf16*>(st_smem + stg * 2 * kStTile);
        bf16* ds = reinterpret_cast<bf16*>(st_smem + stg * 2 * kStTile + kStTile);
        for (int i = threadIdx.x; i < bytes / 16; i += kThreads) {
            float v[VEC], d[VEC];
            load_vec(ys + i * VEC, v);
            load_vec(ds + i * VEC, d);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float da = SE ? fmaf(d[j], m0[j], m1[j]) : d[j] * m0[j];
                const float dz = fmaf(a[j], v[j], b[j]) > 0.f ? da : 0.f;
                d[j] = fmaf(-v[j], k2[j], dz) + k0[j];
            }
            store_vec(ds + i * VEC, d);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store_1d(ob + o, ds, (uint32_t)bytes);
            tma_store_commit();
            // the stage tile t-1 lived in is free once its store has read it: refill it with tile t-1+kStStages
            if (t >= 1) {
                tma_store_wait_read1();
                if (t - 1 + kStStages < ntiles) issue(t - 1 + kStStages);
            } else if (kStStages - 1 < ntiles) {
                issue(kStStages - 1);          // the one stage the prologue left empty
            }
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
}

// The read-only reduction passes as the same shared-memory stream (no store: a stage is refilled as soon as every
// thread has consumed it, so all kRdStages tiles are in flight).
//   MODE 0: (sum m*dO, sum m*dO*(y-mean)) from dO and Y — bn_bwd_reduce without SE in front, and se_bn_bwd_reduce
//   MODE 1: (sum m, sum m*y) from Y alone — se_squeeze                                  m = (a*y+b > 0)
// parts: [(b*gridDim.x + blk)][2][C], as the register kernels write them.
constexpr int kRdSmemBytes = 3 * 2 * kStTile;          // 3 stages of 2 x 8 KB (two operands) or 1 x 16 KB (one)

template <int MODE>
__global__ void __launch_bounds__(kThreads, 3) reduce_stream_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ y,
                                                                    int C, long pix_per_img, long chunk,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ shift,
                                                                    const float* __restrict__ mean, float* __restrict__ parts) {
    constexpr int VEC = 8;
    constexpr int NOP = MODE == 0 ? 2 : 1;
    constexpr int TILE = NOP == 1 ? 2 * kStTile : kStTile;      // one operand: twice the tile, same bytes per stage
    constexpr int kRdStages = kRdSmemBytes / (NOP * TILE);
    extern __shared__ __align__(128) uint8_t st_smem[];        // [stage][y tile | dO tile]
    __shared__ __align__(8) uint64_t full_bar[kRdStages];
    const int vpr = C / VEC;
    const int cv = threadIdx.x & (vpr - 1);
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    const long off0 = ((long)blockIdx.y * pix_per_img + p0) * C * 2;
    const long nbytes = p1 > p0 ? (p1 - p0) * C * 2 : 0;
    const int ntiles = (int)((nbytes + TILE - 1) / TILE);
    const uint8_t* yb = reinterpret_cast<const uint8_t*>(y) + off0;
    const uint8_t* db = MODE == 0 ? reinterpret_cast<const uint8_t*>(dout) + off0 : nullptr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kRdStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int stg = t % kRdStages;
        const long o = (long)t * TILE;
        const uint32_t bytes = (uint32_t)(nbytes - o < TILE ? nbytes - o : TILE);
        mbar_expect_tx(&full_bar[stg], NOP * bytes);
        bulk_load_1d(st_smem + stg * NOP * TILE, yb + o, bytes, &full_bar[stg]);
        if (MODE == 0) bulk_load_1d(st_smem + stg * NOP * TILE + TILE, db + o, bytes, &full_bar[stg]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kRdStages && t < ntiles; ++t) issue(t);
    float a[VEC], b[VEC], mu[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        a[i] = scale[cv * VEC + i]; b[i] = shift[cv * VEC + i];
        mu[i] = MODE == 0 ? mean[cv * VEC + i] : 0.f;
    }
    float acc[2][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
    for (int t = 0; t < ntiles; ++t) {
        const int stg = t % kRdStages;
        const long o = (long)t * TILE;
        const int bytes = (int)(nbytes - o < TILE ? nbytes - o : TILE);
        mbar_wait(&full_bar[stg], (uint32_t)((t / kRdStages) & 1));
        const bf16* ys = reinterpret_cast<const bf16*>(st_smem + stg * NOP * TILE);
        const bf16* ds = reinterpret_cast<const bf16*>(st_smem + stg * NOP * TILE + TILE);
        for (int i = threadIdx.x; i < bytes / 16; i += kThreads) {
            float v[VEC], d[VEC];
            load_vec(ys + i * VEC, v);
            if (MODE == 0) load_vec(ds + i * VEC, d);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const bool m = fmaf(a[j], v[j], b[j]) > 0.f;
                if (MODE == 0) {
                    const float dm = m ? d[j] : 0.f;
                    acc[0][j] += dm;
                    acc[1][j] = fmaf(dm, v[j] - mu[j], acc[1][j]);
                } else if (m) {
                    acc[0][j] += 1.f; acc[1][j] += v[j];
                }
            }
        }
        __syncthreads();                                   // every thread is done with this stage
        if (threadIdx.x == 0 && t + kRdStages < ntiles) issue(t + kRdStages);
    }
    block_reduce_rows<2, VEC>(acc, C, vpr, kThreads / vpr, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}
template <typename K> static int resident_blocks_smem(K kernel, int smem) {
    int v = 0;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kThreads, smem) != cudaSuccess || v < 1) return 0;
    return v * num_sms();
}
// dense bf16 tensors whose channel-vector count is a power of two dividing the block: the shapes the streams take
template <typename T> static inline bool stream_ok(int C, int ld0, int ld1) {
    const int vpr8 = C / 8;
    return g_apply_stream > 0 && sizeof(T) == 2 && ld0 == C && ld1 == C && C % 8 == 0 && vpr8 >= 1 && vpr8 <= kThreads &&
           (vpr8 & (vpr8 - 1)) == 0;
}

__global__ void bn_bwd_finalize_kernel(const float* __restrict__ parts, int nparts, int C, double count,
                                       const float* __restrict__ gamma, const float* __restrict__ invstd,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef) {
    double tot[2];
    if (!sum_parts32<2>(parts, nparts, C, tot)) return;
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const double s = tot[0], q = tot[1] * (double)invstd[c];
    dbeta[c] = (float)s;
    dgamma[c] = (float)q;
    coef[c] = gamma[c] * invstd[c];
    coef[C + c] = (float)(s / count);
    coef[2 * C + c] = (float)(q / count);
}

// ---------------------------------------------------------------------------------------------------------
// SE + ReLU + BN backward, stage 1 when an SE layer follows the block: ONE pass over dO and Y2 instead of the
// se_bwd_reduce + bn_bwd_reduce pair.  With m = (a*y+b > 0), per (image, channel):
//   S1 = sum m*dO,  S2 = sum m*dO*(y-mean)          (this kernel, backward)
//   S3 = sum m,     S4 = sum m*(y-mean)             (se_squeeze_kernel, forward: they do not depend on dO)
// Everything downstream is linear in these: the SE excitation gradient ds = sum dO*relu(a*y+b) = a*S2 + beta*S1
// (beta = a*mean + b), and once the FC chain has produced dp[b,c]:
//   sum dz = sum_b s*S1 + dp/HW*S3,      sum dz*(y-mean) = sum_b s*S2 + dp/HW*S4.
// parts: [(b*nblk + blk)][2][C]
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 4) se_bn_bwd_reduce_kernel(const T* __restrict__ dout, int ldd,
                                                                       const T* __restrict__ y, int ldy, int C,
                                                                       long pix_per_img, long chunk,
                                                                       const float* __restrict__ scale,
                                                                       const float* __restrict__ shift,
                                                                       const float* __restrict__ mean,
                                                                       float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long base = (long)blockIdx.y * pix_per_img;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    float acc[2][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
    if (r < rows) {
        float a[VEC], b[VEC], mu[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { a[i] = scale[cv * VEC + i]; b[i] = shift[cv * VEC + i]; mu[i] = mean[cv * VEC + i]; }
#pragma unroll 4
        for (long p = p0 + r; p < p1; p += rows) {
            float v[VEC], d[VEC];
            load_vec(y + (base + p) * ldy + cv * VEC, v);
            load_vec(dout + (base + p) * ldd + cv * VEC, d);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float dm = fmaf(a[i], v[i], b[i]) > 0.f ? d[i] : 0.f;
                acc[0][i] += dm;
                acc[1][i] = fmaf(dm, v[i] - mu[i], acc[1][i]);
            }
        }
    }
    block_reduce_rows<2, VEC>(acc, C, vpr, rows, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// ---------------------------------------------------------------------------------------------------------
// The SE + ReLU + BN backward of an ENCODER block, with the gradient of its output formed on the fly.  That output
// feeds the decoder (skip gradient sg, full resolution) and the next level's MaxPool2d(2) (gradient dpooled routed to
// the window position `pos`): dO[p] = sg[p] + (pos[quad(p)] == k(p) ? dpooled[quad(p)] : 0).  Materialising dO first
// (pool_bwd_add) costs a 2.25*N pass; here both passes that consume it (reduction and apply) rebuild it per 2x2 quad
// from one dpooled vector and one 8-byte position word.  One thread = a run of quads x one channel vector;
// grid = (blocks per image, B).  APPLY = false: parts[(b*nblk+blk)][2][C] = (sum m*dO, sum m*dO*(y-mean));
// APPLY = true: dY as in bn_bwd_kernel<T, true, true>.
// ---------------------------------------------------------------------------------------------------------
template <typename T, bool APPLY>
__global__ void __launch_bounds__(kThreads, 2) se_bn_bwd_pool_kernel(const T* __restrict__ sg, int lds,
                                                                     const T* __restrict__ dpooled, int ldp,
                                                                     const uint8_t* __restrict__ pos,
                                                                     const T* __restrict__ y, int ldy, T* __restrict__ dy,
                                                                     int lddy, int H, int W, int C, long chunk, float inv_hw,
                                                                     const float* __restrict__ scale,
                                                                     const float* __restrict__ shift,
                                                                     const float* __restrict__ mean,
                                                                     const float* __restrict__ invstd,
                                                                     const float* __restrict__ s, const float* __restrict__ dp,
                                                                     const float* __restrict__ coef,
                                                                     float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const int Ho = H >> 1, Wo = W >> 1;
    const long nquad = (long)Ho * Wo;
    const int b = blockIdx.y;
    const long q0 = (long)blockIdx.x * chunk;
    long q1 = q0 + chunk; if (q1 > nquad) q1 = nquad;
    float acc[2][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
    if (r < rows) {
        float a[VEC], bb[VEC], m0[VEC], m1[VEC], k2[VEC], k0[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const int c = cv * VEC + i;
            a[i] = scale[c]; bb[i] = shift[c];
            if (APPLY) {
                const float g = coef[c], c1 = coef[C + c], c2 = coef[2 * C + c];
                k2[i] = g * c2 * invstd[c];
                k0[i] = mean[c] * k2[i] - g * c1;
                m0[i] = g * s[(long)b * C + c]; m1[i] = g * dp[(long)b * C + c] * inv_hw;
            } else {
                k2[i] = mean[c]; k0[i] = 0.f; m0[i] = 0.f; m1[i] = 0.f;
            }
        }
#pragma unroll 2
        for (long q = q0 + r; q < q1; q += rows) {
            const int wo = (int)(q % Wo), ho = (int)(q / Wo);
            const long qg = (long)b * nquad + q;
            float g[VEC];
            load_vec(dpooled + qg * ldp + cv * VEC, g);
            uint8_t code[VEC];
            const uint8_t* src = pos + qg * C + cv * VEC;
            if (VEC == 8) {
                const uint2 t = *reinterpret_cast<const uint2*>(src);
#pragma unroll
                for (int i = 0; i < 4; ++i) { code[i] = (t.x >> (8 * i)) & 0xff; code[(4 + i) % VEC] = (t.y >> (8 * i)) & 0xff; }
            } else {
                const uint32_t t = *reinterpret_cast<const uint32_t*>(src);
#pragma unroll
                for (int i = 0; i < 4; ++i) code[i] = (t >> (8 * i)) & 0xff;
            }
            const long pbase = ((long)b * H + 2 * ho) * W + 2 * wo;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long p = pbase + (k >> 1) * W + (k & 1);
                float v[VEC], d[VEC];
                load_vec(y + p * ldy + cv * VEC, v);
                load_vec(sg + p * lds + cv * VEC, d);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float dO = code[i] == k ? d[i] + g[i] : d[i];
                    const bool on = fmaf(a[i], v[i], bb[i]) > 0.f;
                    if (APPLY) {
                        const float dz = on ? fmaf(dO, m0[i], m1[i]) : 0.f;
                        d[i] = fmaf(-v[i], k2[i], dz) + k0[i];
                    } else {
                        const float dm = on ? dO : 0.f;
                        acc[0][i] += dm;
                        acc[1][i] = fmaf(dm, v[i] - k2[i], acc[1][i]);
                    }
                }
                if (APPLY) store_vec(dy + p * lddy + cv * VEC, d);
            }
        }
    }
    if (!APPLY)
        block_reduce_rows<2, VEC>(acc, C, vpr, rows, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// The same two passes as a shared-memory stream (bf16, C = 64 / 128 / 256): a tile is one quad-row segment of
// QW = 2048/C quads — sg and Y as [2 rows][2*QW px][C] boxes, dpooled and pos as [QW][C] — fetched by four TMA tensor
// copies (so sg may be a channel slice of the decoder's concat gradient), one quad x 8 channels per thread.  APPLY
// writes dY into one of two staging tiles and stores it with one TMA copy, so the input stages refill at once.
// Same arithmetic as above.
constexpr int kQpStages = 2;
constexpr int kQpTileSg = 16384, kQpTileDp = 4096, kQpTilePos = 2048;
constexpr int kQpStageBytes = 2 * kQpTileSg + kQpTileDp + kQpTilePos;
constexpr int kQpSmemBytes = kQpStages * kQpStageBytes + 1024;
constexpr int kQpSmemBytesApply = kQpSmemBytes + 2 * kQpTileSg;       // + two dY staging tiles
struct QpMaps { CUtensorMap sg, y, dp, pos, dy; };

template <bool APPLY>
__global__ void __launch_bounds__(kThreads, 2) se_bn_bwd_pool_stream_kernel(const __grid_constant__ QpMaps maps, int H, int W, int C,
                                                                            int tiles_per_block, float inv_hw,
                                                                            const float* __restrict__ scale,
                                                                            const float* __restrict__ shift,
                                                                            const float* __restrict__ mean,
                                                                            const float* __restrict__ invstd,
                                                                            const float* __restrict__ s, const float* __restrict__ dp,
                                                                            const float* __restrict__ coef, float* __restrict__ parts) {
    extern __shared__ uint8_t qp_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(qp_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kQpStages];
    constexpr int VEC = 8;
    const int vpr = C / VEC, QW = kThreads / vpr;
    const int ql = threadIdx.x / vpr, cv = threadIdx.x % vpr;          // quad within the tile, channel vector
    const int Ho = H >> 1, Wo = W >> 1;
    const int nseg = (Wo + QW - 1) / QW;
    const int ntile_img = Ho * nseg;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * tiles_per_block;
    int t1 = t0 + tiles_per_block; if (t1 > ntile_img) t1 = ntile_img;
    const int ntiles = t1 > t0 ? t1 - t0 : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kQpStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        uint8_t* st = smem + (t % kQpStages) * kQpStageBytes;
        const int tt = t0 + t, ho = tt / nseg, w0 = (tt % nseg) * QW;
        mbar_expect_tx(&full_bar[t % kQpStages], kQpStageBytes);
        tma_load_4d(&maps.sg, &full_bar[t % kQpStages], st, 0, 2 * w0, 2 * ho, b);
        tma_load_4d(&maps.y, &full_bar[t % kQpStages], st + kQpTileSg, 0, 2 * w0, 2 * ho, b);
        tma_load_4d(&maps.dp, &full_bar[t % kQpStages], st + 2 * kQpTileSg, 0, w0, ho, b);
        tma_load_4d(&maps.pos, &full_bar[t % kQpStages], st + 2 * kQpTileSg + kQpTileDp, 0, w0, ho, b);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kQpStages && t < ntiles; ++t) issue(t);
    uint8_t* out_tiles = smem + kQpStages * kQpStageBytes;             // APPLY: dY staging, double-buffered
    float a[VEC], bb[VEC], m0[VEC], m1[VEC], k2[VEC], k0[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = cv * VEC + i;
        a[i] = scale[c]; bb[i] = shift[c];
        if (APPLY) {
            const float g = coef[c], c1 = coef[C + c], c2 = coef[2 * C + c];
            k2[i] = g * c2 * invstd[c];
            k0[i] = mean[c] * k2[i] - g * c1;
            m0[i] = g * s[(long)b * C + c]; m1[i] = g * dp[(long)b * C + c] * inv_hw;
        } else {
            k2[i] = mean[c]; k0[i] = 0.f; m0[i] = 0.f; m1[i] = 0.f;
        }
    }
    float acc[2][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
    for (int t = 0; t < ntiles; ++t) {
        uint8_t* st = smem + (t % kQpStages) * kQpStageBytes;
        mbar_wait(&full_bar[t % kQpStages], (uint32_t)((t / kQpStages) & 1));
        const bf16* sgs = reinterpret_cast<const bf16*>(st);
        const bf16* ys = reinterpret_cast<const bf16*>(st + kQpTileSg);
        bf16* outs = reinterpret_cast<bf16*>(out_tiles + (t & 1) * kQpTileSg);
        float g[VEC];
        load_vec(reinterpret_cast<const bf16*>(st + 2 * kQpTileSg) + ((long)ql * C + cv * VEC), g);
        const uint2 tp = *reinterpret_cast<const uint2*>(st + 2 * kQpTileSg + kQpTileDp + ql * C + cv * VEC);
        uint8_t code[VEC];
#pragma unroll
        for (int i = 0; i < 4; ++i) { code[i] = (tp.x >> (8 * i)) & 0xff; code[4 + i] = (tp.y >> (8 * i)) & 0xff; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long e = ((long)((k >> 1) * 2 * QW + 2 * ql + (k & 1))) * C + cv * VEC;      // [row][px][C]
            float v[VEC], d[VEC];
            load_vec(ys + e, v);
            load_vec(sgs + e, d);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float dO = code[i] == k ? d[i] + g[i] : d[i];
                const bool on = fmaf(a[i], v[i], bb[i]) > 0.f;
                if (APPLY) {
                    const float dz = on ? fmaf(dO, m0[i], m1[i]) : 0.f;
                    d[i] = fmaf(-v[i], k2[i], dz) + k0[i];
                } else {
                    const float dm = on ? dO : 0.f;
                    acc[0][i] += dm;
                    acc[1][i] = fmaf(dm, v[i] - k2[i], acc[1][i]);
                }
            }
            if (APPLY) store_vec(outs + e, d);
        }
        if (APPLY) {
            // the other staging tile (written next iteration) must have been read by its store, issued one tile ago
            if (threadIdx.x == 0) tma_store_wait_read();
            fence_proxy_async_smem();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (APPLY) {
                const int tt = t0 + t, ho = tt / nseg, w0 = (tt % nseg) * QW;
                tma_store_4d(&maps.dy, outs, 0, 2 * w0, 2 * ho, b);
                tma_store_commit();
            }
            if (t + kQpStages < ntiles) issue(t + kQpStages);      // the input stage is free as soon as everyone has read it
        }
    }
    if (APPLY) { if (threadIdx.x == 0) tma_store_wait_all(); }
    else block_reduce_rows<2, VEC>(acc, C, vpr, kThreads / vpr, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// tensor maps + launch geometry of se_bn_bwd_pool_stream_kernel; returns 0 when the shape is not streamable
static int qp_setup(QpMaps* m, const void* sg, int lds, const void* dpooled, int ldp, const uint8_t* pos, const void* y, int ldy,
                    void* dy, int lddy, int B, int H, int W, int C) {
    const int QW = kThreads / (C / 8);
    int rc;
    if ((rc = make_tmap_nhwc(&m->sg, sg, 2, C, W, H, B, lds, 2 * QW, 2)) < 0) return rc;
    if ((rc = make_tmap_nhwc(&m->y, y, 2, C, W, H, B, ldy, 2 * QW, 2)) < 0) return rc;
    if ((rc = make_tmap_nhwc(&m->dp, dpooled, 2, C, W / 2, H / 2, B, ldp, QW, 1)) < 0) return rc;
    if ((rc = make_tmap_nhwc(&m->pos, pos, 1, C, W / 2, H / 2, B, C, QW, 1)) < 0) return rc;
    if (dy && (rc = make_tmap_nhwc(&m->dy, dy, 2, C, W, H, B, lddy, 2 * QW, 2)) < 0) return rc;
    return 1;
}
template <typename T> static inline bool qp_ok(int C, int lds, int ldp, int ldy, int lddy) {
    return g_apply_stream > 0 && sizeof(T) == 2 && (C == 64 || C == 128 || C == 256) && lds % 8 == 0 && ldp % 8 == 0 &&
           ldy % 8 == 0 && lddy % 8 == 0;
}

// ---------------------------------------------------------------------------------------------------------
// Block output o = relu(a*y+b) * s[b,c] (+ fused MaxPool2d(2)) as shared-memory streams (bf16).
//   se_scale_pool_stream_kernel: quad-row tiles as above — Y in by one TMA tensor copy; o (possibly a channel slice of
//     the decoder's concat buffer), pooled and pos out by three, from double-buffered staging (C = 64 / 128 / 256)
//   se_scale_stream_kernel: dense 1-D tiles, transformed in place (any power-of-two channel-vector count)
// Same arithmetic, rounding and tie-break as se_scale_pool_kernel / se_scale_kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSpStages = 3;
constexpr int kSpOutBytes = kQpTileSg + kQpTileDp + kQpTilePos;                      // o | pooled | pos
constexpr int kSpSmemBytes = kSpStages * kQpTileSg + 2 * kSpOutBytes + 1024;
struct SpMaps { CUtensorMap y, out, pooled, pos; };

__global__ void __launch_bounds__(kThreads, 2) se_scale_pool_stream_kernel(const __grid_constant__ SpMaps maps, int H, int W, int C,
                                                                           int tiles_per_block, const float* __restrict__ scale,
                                                                           const float* __restrict__ shift,
                                                                           const float* __restrict__ s) {
    extern __shared__ uint8_t qp_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(qp_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kSpStages];
    constexpr int VEC = 8;
    const int vpr = C / VEC, QW = kThreads / vpr;
    const int ql = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const int Ho = H >> 1, Wo = W >> 1;
    const int nseg = (Wo + QW - 1) / QW;
    const int ntile_img = Ho * nseg;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * tiles_per_block;
    int t1 = t0 + tiles_per_block; if (t1 > ntile_img) t1 = ntile_img;
    const int ntiles = t1 > t0 ? t1 - t0 : 0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kSpStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const int tt = t0 + t, ho = tt / nseg, w0 = (tt % nseg) * QW;
        mbar_expect_tx(&full_bar[t % kSpStages], kQpTileSg);
        tma_load_4d(&maps.y, &full_bar[t % kSpStages], smem + (t % kSpStages) * kQpTileSg, 0, 2 * w0, 2 * ho, b);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kSpStages && t < ntiles; ++t) issue(t);
    uint8_t* out_tiles = smem + kSpStages * kQpTileSg;
    float a[VEC], sh[VEC], g[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        a[i] = scale[cv * VEC + i];
        sh[i] = shift[cv * VEC + i];
        g[i] = s ? s[(long)b * C + cv * VEC + i] : 1.f;
    }
    for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&full_bar[t % kSpStages], (uint32_t)((t / kSpStages) & 1));
        const bf16* ys = reinterpret_cast<const bf16*>(smem + (t % kSpStages) * kQpTileSg);
        uint8_t* ob = out_tiles + (t & 1) * kSpOutBytes;
        bf16* os = reinterpret_cast<bf16*>(ob);
        float best[VEC];
        uint8_t code[VEC];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long e = ((long)((k >> 1) * 2 * QW + 2 * ql + (k & 1))) * C + cv * VEC;
            float v[VEC];
            load_vec(ys + e, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float tv = fmaxf(fmaf(a[i], v[i], sh[i]), 0.f) * g[i];
                tv = round_to(tv, (const bf16*)nullptr);
                v[i] = tv;
                if (k == 0) { best[i] = tv; code[i] = 0; }
                else if (tv > best[i] || tv != tv) { best[i] = tv; code[i] = (uint8_t)k; }
            }
            store_vec(os + e, v);
        }
        store_vec(reinterpret_cast<bf16*>(ob + kQpTileSg) + ((long)ql * C + cv * VEC), best);
        uint2 tp;
        tp.x = code[0] | (code[1] << 8) | (code[2] << 16) | (code[3] << 24);
        tp.y = code[4] | (code[5] << 8) | (code[6] << 16) | (code[7] << 24);
        *reinterpret_cast<uint2*>(ob + kQpTileSg + kQpTileDp + ql * C + cv * VEC) = tp;
        if (threadIdx.x == 0) tma_store_wait_read();        // the other staging buffer's stores (one tile ago) have read it
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int tt = t0 + t, ho = tt / nseg, w0 = (tt % nseg) * QW;
            tma_store_4d(&maps.out, ob, 0, 2 * w0, 2 * ho, b);
            tma_store_4d(&maps.pooled, ob + kQpTileSg, 0, w0, ho, b);
            tma_store_4d(&maps.pos, ob + kQpTileSg + kQpTileDp, 0, w0, ho, b);
            tma_store_commit();
            if (t + kSpStages < ntiles) issue(t + kSpStages);
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
}

constexpr int kScTile = 16384;
constexpr int kScStages = 4;
constexpr int kScSmemBytes = kScStages * kScTile;
__global__ void __launch_bounds__(kThreads, 3) se_scale_stream_kernel(const bf16* __restrict__ y, bf16* __restrict__ out, int C,
                                                                      long pix_per_img, long chunk,
                                                                      const float* __restrict__ scale,
                                                                      const float* __restrict__ shift,
                                                                      const float* __restrict__ s) {
    extern __shared__ __align__(128) uint8_t st_smem[];
    __shared__ __align__(8) uint64_t full_bar[kScStages];
    constexpr int VEC = 8;
    const int vpr = C / VEC;
    const int cv = threadIdx.x & (vpr - 1);
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    const long off0 = ((long)blockIdx.y * pix_per_img + p0) * C * 2;
    const long nbytes = (p1 - p0) * C * 2;
    const int ntiles = (int)((nbytes + kScTile - 1) / kScTile);
    const uint8_t* yb = reinterpret_cast<const uint8_t*>(y) + off0;
    uint8_t* ob = reinterpret_cast<uint8_t*>(out) + off0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kScStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const long o = (long)t * kScTile;
        const uint32_t bytes = (uint32_t)(nbytes - o < kScTile ? nbytes - o : kScTile);
        mbar_expect_tx(&full_bar[t % kScStages], bytes);
        bulk_load_1d(st_smem + (t % kScStages) * kScTile, yb + o, bytes, &full_bar[t % kScStages]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kScStages - 1 && t < ntiles; ++t) issue(t);
    float a[VEC], sh[VEC], g[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        a[i] = scale[cv * VEC + i];
        sh[i] = shift[cv * VEC + i];
        g[i] = s ? s[(long)blockIdx.y * C + cv * VEC + i] : 1.f;
    }
    for (int t = 0; t < ntiles; ++t) {
        const long o = (long)t * kScTile;
        const int bytes = (int)(nbytes - o < kScTile ? nbytes - o : kScTile);
        mbar_wait(&full_bar[t % kScStages], (uint32_t)((t / kScStages) & 1));
        bf16* ys = reinterpret_cast<bf16*>(st_smem + (t % kScStages) * kScTile);
        for (int i = threadIdx.x; i < bytes / 16; i += kThreads) {
            float v[VEC];
            load_vec(ys + i * VEC, v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) v[j] = fmaxf(fmaf(a[j], v[j], sh[j]), 0.f) * g[j];
            store_vec(ys + i * VEC, v);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store_1d(ob + o, ys, (uint32_t)bytes);
            tma_store_commit();
            if (t >= 1) {
                tma_store_wait_read1();
                if (t - 1 + kScStages < ntiles) issue(t - 1 + kScStages);
            } else if (kScStages - 1 < ntiles) {
                issue(kScStages - 1);
            }
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
}

// SE squeeze (forward): per (image, channel) S3 = sum m and Sy = sum m*y with m = (a*y+b > 0).  The squeeze itself
// follows from them, sum relu(a*y+b) = a*Sy + b*S3, and so does the centred sum the backward pass needs,
// S4 = Sy - mean*S3 (both formed in double by se_fc_kernel<3>).  Two predicated adds per element keep this read-only
// pass memory-bound.  parts: [(b*nblk + blk)][2][C]
template <typename T>
__global__ void __launch_bounds__(kThreads, 4) se_squeeze_kernel(const T* __restrict__ y, int ldy, int C,
                                                                 long pix_per_img, long chunk,
                                                                 const float* __restrict__ scale,
                                                                 const float* __restrict__ shift,
                                                                 float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long base = (long)blockIdx.y * pix_per_img;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > pix_per_img) p1 = pix_per_img;
    float acc[2][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = acc[1][i] = 0.f;
    if (r < rows) {
        float a[VEC], b[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) { a[i] = scale[cv * VEC + i]; b[i] = shift[cv * VEC + i]; }
#pragma unroll 8
        for (long p = p0 + r; p < p1; p += rows) {
            float v[VEC];
            load_vec(y + (base + p) * ldy + cv * VEC, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                if (fmaf(a[i], v[i], b[i]) > 0.f) { acc[0][i] += 1.f; acc[1][i] += v[i]; }
            }
        }
    }
    block_reduce_rows<2, VEC>(acc, C, vpr, rows, parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// SE backward FC chain fed by the four merged sums (one block per image); also stores the per-image sums
// sums[b][4][C] for bn_bwd_finalize_se_kernel.
__global__ void __launch_bounds__(256) se_fc_bwd_fused_kernel(const float* __restrict__ parts, int nparts, int img_rows, int C, int Cr,
                                                              const float* __restrict__ w1, const float* __restrict__ w2,
                                                              const float* __restrict__ z, const float* __restrict__ s,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ sums34,
                                                              float* __restrict__ sums,
                                                              float* __restrict__ dpre2_out, float* __restrict__ dz_out,
                                                              float* __restrict__ dp_out) {
    extern __shared__ float sm[];
    float* dpre2 = sm;       // [C]
    float* dz = sm + C;      // [Cr]
    const int b = blockIdx.x, tid = threadIdx.x;
    for (int c = tid; c < C; c += blockDim.x) {
        double t[2] = {0.0, 0.0};
        for (int i = 0; i < nparts; ++i) {
            const float* row = parts + ((long)b * img_rows + i) * 2 * C + c;
            t[0] += (double)row[0];
            t[1] += (double)row[C];
        }
        sums[((long)b * 4 + 0) * C + c] = (float)t[0];
        sums[((long)b * 4 + 1) * C + c] = (float)t[1];
        sums[((long)b * 4 + 2) * C + c] = sums34[((long)b * 2 + 0) * C + c];
        sums[((long)b * 4 + 3) * C + c] = sums34[((long)b * 2 + 1) * C + c];
        const double a = scale[c], beta = (double)scale[c] * (double)mean[c] + (double)shift[c];
        const float ds = (float)(a * t[1] + beta * t[0]);
        const float sv = s[(long)b * C + c];
        const float v = ds * sv * (1.f - sv);
        dpre2[c] = v;
        dpre2_out[(long)b * C + c] = v;
    }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    // dz[j] = (sum_c dpre2[c] * W2[c][j]) * (z[j] > 0).  W2 is (C, Cr) row-major: lanes run along j (a warp reads whole
    // rows, coalesced) and each warp walks a contiguous slice of c; the warps' partial vectors are then added in a fixed
    // order.  (Lanes along c read one element out of every row: 32 sectors per load, 60 us per call at C = 1024.)
    float* wpart = sm + C + Cr;                       // [nwarps][Cr]
    const int cper = (C + nwarps - 1) / nwarps;
    for (int j0 = 0; j0 < Cr; j0 += 32) {
        const int j = j0 + lane;
        if (j < Cr) {
            float t = 0.f;
            const int c0 = warp * cper, c1 = c0 + cper < C ? c0 + cper : C;
            for (int c = c0; c < c1; ++c) t = fmaf(dpre2[c], __ldg(w2 + (long)c * Cr + j), t);
            wpart[warp * Cr + j] = t;
        }
    }
    __syncthreads();
    for (int j = tid; j < Cr; j += blockDim.x) {
        float t = 0.f;
        for (int w = 0; w < nwarps; ++w) t += wpart[w * Cr + j];
        t = z[(long)b * Cr + j] > 0.f ? t : 0.f;
        dz[j] = t;
        dz_out[(long)b * Cr + j] = t;
    }
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {
        float t = 0.f;
        for (int j = 0; j < Cr; ++j) t = fmaf(dz[j], w1[(long)j * C + c], t);
        dp_out[(long)b * C + c] = t;
    }
}

// BN backward finalize from the per-image merged sums: 32 channels per block, 8 image lanes folded in double.
__global__ void bn_bwd_finalize_se_kernel(const float* __restrict__ sums, int B, int C, double count, double inv_hw,
                                          const float* __restrict__ gamma, const float* __restrict__ invstd,
                                          const float* __restrict__ s, const float* __restrict__ dp,
                                          float* __restrict__ dgamma, float* __restrict__ dbeta,
                                          float* __restrict__ coef) {
    __shared__ double red[2][8][32];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double a0 = 0.0, a1 = 0.0;
    if (c < C) {
        for (int b = pl; b < B; b += 8) {
            const float* row = sums + (long)b * 4 * C + c;
            const double sv = s[(long)b * C + c], dpv = (double)dp[(long)b * C + c] * inv_hw;
            a0 += sv * (double)row[0] + dpv * (double)row[2 * (long)C];
            a1 += sv * (double)row[(long)C] + dpv * (double)row[3 * (long)C];
        }
    }
    red[0][pl][cl] = a0; red[1][pl][cl] = a1;
    __syncthreads();
    if (pl != 0 || c >= C) return;
    double sdz = 0.0, q = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sdz += red[0][k][cl]; q += red[1][k][cl]; }
    q *= (double)invstd[c];
    dbeta[c] = (float)sdz;
    dgamma[c] = (float)q;
    coef[c] = gamma[c] * invstd[c];
    coef[C + c] = (float)(sdz / count);
    coef[2 * C + c] = (float)(q / count);
}

// ---------------------------------------------------------------------------------------------------------
// Bilinear resize guard of the decoder (UCA:138-157): torchvision F_T.resize(tensor, size, BILINEAR) ==
// interpolate(mode='bilinear', align_corners=False, antialias=True).  It only ever up-samples here
// (2*floor(h/2) -> h), where the antialias triangle filter reduces to the ordinary two taps with index clamping
// (the two-tap form is pinned against torch.nn.functional.interpolate(antialias=True) by the CPU tests).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_tap(int o, int n_in, float scale, int& i0, int& i1, float& lam) {
    const float src = fmaxf(((float)o + 0.5f) * scale - 0.5f, 0.f);
    i0 = min((int)src, n_in - 1);
    i1 = min(i0 + 1, n_in - 1);
    lam = src - (float)i0;
}

// one thread = one output pixel x one channel vector
template <typename T>
__global__ void __launch_bounds__(kThreads) resize_bilinear_fwd_kernel(const T* __restrict__ x, int ldx, int h, int w,
                                                                       T* __restrict__ out, int ldo, int H, int W, int B,
                                                                       int C, float sh, float sw) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC;
    const long gid = (long)blockIdx.x * kThreads + threadIdx.x;
    const long pix = gid / vpr;
    const int cv = (int)(gid % vpr);
    if (pix >= (long)B * H * W) return;
    const int ow = (int)(pix % W), oh = (int)((pix / W) % H);
    const long b = pix / ((long)W * H);
    int h0, h1, w0, w1; float lh, lw;
    bilinear_tap(oh, h, sh, h0, h1, lh);
    bilinear_tap(ow, w, sw, w0, w1, lw);
    float v00[VEC], v01[VEC], v10[VEC], v11[VEC];
    const T* xb = x + b * h * w * (long)ldx + cv * VEC;
    load_vec(xb + ((long)h0 * w + w0) * ldx, v00);
    load_vec(xb + ((long)h0 * w + w1) * ldx, v01);
    load_vec(xb + ((long)h1 * w + w0) * ldx, v10);
    load_vec(xb + ((long)h1 * w + w1) * ldx, v11);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float top = v00[i] * (1.f - lw) + v01[i] * lw;
        const float bot = v10[i] * (1.f - lw) + v11[i] * lw;
        v00[i] = top * (1.f - lh) + bot * lh;
    }
    store_vec(out + pix * ldo + cv * VEC, v00);
}

// adjoint, as a gather: one thread = one INPUT pixel x one channel vector; it visits the few output pixels whose
// taps can reference it (deterministic, no atomics).
template <typename T>
__global__ void __launch_bounds__(kThreads) resize_bilinear_bwd_kernel(const T* __restrict__ dout, int ldd, int H, int W,
                                                                       T* __restrict__ dx, int ldx, int h, int w, int B,
                                                                       int C, float sh, float sw) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC;
    const long gid = (long)blockIdx.x * kThreads + threadIdx.x;
    const long pix = gid / vpr;
    const int cv = (int)(gid % vpr);
    if (pix >= (long)B * h * w) return;
    const int iw = (int)(pix % w), ih = (int)((pix / w) % h);
    const long b = pix / ((long)w * h);
    // output rows / columns whose source coordinate can fall in (i-1, i+1)
    const int oh_lo = max(0, (int)floorf(((float)ih - 0.5f) / sh - 0.5f) - 1);
    const int oh_hi = min(H - 1, (int)ceilf(((float)ih + 1.5f) / sh - 0.5f) + 1);
    const int ow_lo = max(0, (int)floorf(((float)iw - 0.5f) / sw - 0.5f) - 1);
    const int ow_hi = min(W - 1, (int)ceilf(((float)iw + 1.5f) / sw - 0.5f) + 1);
    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;
    const T* db = dout + b * H * W * (long)ldd + cv * VEC;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
        int h0, h1; float lh;
        bilinear_tap(oh, h, sh, h0, h1, lh);
        const float wh = (h0 == ih ? 1.f - lh : 0.f) + (h1 == ih ? lh : 0.f);
        if (wh == 0.f) continue;
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
            int w0, w1; float lw;
            bilinear_tap(ow, w, sw, w0, w1, lw);
            const float ww = (w0 == iw ? 1.f - lw : 0.f) + (w1 == iw ? lw : 0.f);
            if (ww == 0.f) continue;
            float d[VEC];
            load_vec(db + ((long)oh * W + ow) * ldd, d);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[i] = fmaf(wh * ww, d[i], acc[i]);
        }
    }
    store_vec(dx + pix * ldx + cv * VEC, acc);
}

// last row / column of an odd-sized map: MaxPool2d(2) floors, so those pixels only carry the skip gradient
template <typename T>
__global__ void __launch_bounds__(kThreads) pool_bwd_border_kernel(const T* __restrict__ skip_grad, int lds,
                                                                   T* __restrict__ dx, int ldx, int B, int H, int W, int C) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC;
    const int nrow = (H & 1) ? W : 0, ncol = (W & 1) ? (H - (H & 1)) : 0;
    const long gid = (long)blockIdx.x * kThreads + threadIdx.x;
    const long k = gid / vpr;
    const int cv = (int)(gid % vpr);
    if (k >= (long)B * (nrow + ncol)) return;
    const int j = (int)(k % (nrow + ncol));
    const long b = k / (nrow + ncol);
    const int hh = j < nrow ? H - 1 : j - nrow, ww = j < nrow ? j : W - 1;
    const long p = (b * H + hh) * W + ww;
    float v[VEC];
    if (skip_grad) load_vec(skip_grad + p * lds + cv * VEC, v);
    else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = 0.f;
    }
    store_vec(dx + p * ldx + cv * VEC, v);
}

// per-channel sum over pixels (ConvTranspose2d bias gradient)
template <typename T>
__global__ void __launch_bounds__(kThreads) chan_sum_kernel(const T* __restrict__ x, int ld, int C, long npix,
                                                            long chunk, float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > npix) p1 = npix;
    float acc[1][VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[0][i] = 0.f;
    if (r < rows) {
#pragma unroll 4
        for (long p = p0 + r; p < p1; p += rows) {
            float v[VEC];
            load_vec(x + p * ld + cv * VEC, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[0][i] += v[i];
        }
    }
    block_reduce_rows<1, VEC>(acc, C, vpr, rows, parts + (long)blockIdx.x * C);
}
// out[i] = scale * sum_parts parts[p][i]      (generic deterministic second stage)
__global__ void sum_parts_kernel(const float* __restrict__ parts, int nparts, int n, const float* __restrict__ scale_ptr,
                                 float* __restrict__ out) {
    double tot[1];
    if (!sum_parts32<1>(parts, nparts, n, tot)) return;
    const int i = blockIdx.x * 32 + (threadIdx.x & 31);
    out[i] = (float)(tot[0] * (scale_ptr ? (double)*scale_ptr : 1.0));
}

// out[i] = sum over rows of parts[row * row_stride + i], i < n   (double accumulation, fixed order)
__global__ void sum_rows_kernel(const float* __restrict__ parts, int nrows, long row_stride, int n, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = 0.0;
    for (int r = 0; r < nrows; ++r) t += (double)parts[(long)r * row_stride + i];
    out[i] = (float)t;
}

// ---------------------------------------------------------------------------------------------------------
// outc: 1x1 conv C -> NC class logits, NHWC T in, NCHW fp32 out.                       UCA:125,162
// 8 lanes share one pixel (16 B each), shuffle-reduce, then the warp's 32 pixels are written coalesced.
// ---------------------------------------------------------------------------------------------------------
template <typename T, int NC>
__global__ void __launch_bounds__(kThreads) outc_fwd_kernel(const T* __restrict__ x, int ldx, int C,
                                                            const float* __restrict__ w, const float* __restrict__ bias,
                                                            int nc, float* __restrict__ logits, long npix, long HW) {
    constexpr int VEC = VecTraits<T>::N;
    extern __shared__ float wsm[];   // [NC][C]
    for (int i = threadIdx.x; i < NC * C; i += kThreads) wsm[i] = (i / C) < nc ? w[i] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
    const long warp_stride = (long)gridDim.x * (kThreads / 32) * 32;
    const bool one_pass = C == 8 * VEC;        // every lane group covers the pixel in one load: keep its weights in registers
    float wr[NC][VEC];
#pragma unroll
    for (int o = 0; o < NC; ++o)
#pragma unroll
        for (int i = 0; i < VEC; ++i) wr[o][i] = one_pass ? wsm[o * C + sub * VEC + i] : 0.f;
    // each warp walks 32-pixel groups with a grid stride: the weight prologue is paid once per block, not per 256 pixels
    for (long warp_pix0 = ((long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * 32; warp_pix0 < npix;
         warp_pix0 += warp_stride) {
    float keep[NC];
#pragma unroll
    for (int o = 0; o < NC; ++o) keep[o] = 0.f;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const long p = warp_pix0 + it * 4 + grp;
        float acc[NC];
#pragma unroll
        for (int o = 0; o < NC; ++o) acc[o] = 0.f;
        if (p < npix && one_pass) {
            float v[VEC];
            load_vec(x + p * ldx + sub * VEC, v);
#pragma unroll
            for (int o = 0; o < NC; ++o)
#pragma unroll
                for (int i = 0; i < VEC; ++i) acc[o] = fmaf(v[i], wr[o][i], acc[o]);
        } else if (p < npix) {
            for (int c0 = sub * VEC; c0 < C; c0 += 8 * VEC) {
                float v[VEC];
                load_vec(x + p * ldx + c0, v);
#pragma unroll
                for (int o = 0; o < NC; ++o)
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[o] = fmaf(v[i], wsm[o * C + c0 + i], acc[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < NC; ++o) {
            float t = acc[o];
            t += __shfl_xor_sync(0xffffffffu, t, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 2);
            t += __shfl_xor_sync(0xffffffffu, t, 4);
            // lane L finally owns pixel warp_pix0 + L = it*4 + grp  ->  it = L/4, grp = L%4
            const float got = __shfl_sync(0xffffffffu, t, (lane & 3) * 8);
            if ((lane >> 2) == it) keep[o] = got;
        }
    }
    const long p = warp_pix0 + lane;
    if (p < npix) {
        const long b = p / HW, hw = p % HW;
#pragma unroll
        for (int o = 0; o < NC; ++o)
            if (o < nc) logits[(b * nc + o) * HW + hw] = keep[o] + bias[o];
    }
    }
}

// outc backward: g (B,NC,H,W) fp32 = un-normalised dlogits, *gscale = 1/N_valid * upstream grad.
//   dx[p,c] = gscale * sum_o g[o,p] W[o,c]            (NHWC T)
//   parts[blk][o][c] = sum_p g[o,p] x[p,c],  parts[blk][NC*C + o] = sum_p g[o,p]   (finalize multiplies by gscale)
template <typename T, int NC>
__global__ void __launch_bounds__(kThreads) outc_bwd_kernel(const float* __restrict__ g, const float* __restrict__ gscale,
                                                            const T* __restrict__ x, int ldx, T* __restrict__ dx, int lddx,
                                                            int C, const float* __restrict__ w, int nc, long npix, long HW,
                                                            long chunk, float* __restrict__ parts) {
    constexpr int VEC = VecTraits<T>::N;
    const int vpr = C / VEC, rows = kThreads / vpr;
    const int r = threadIdx.x / vpr, cv = threadIdx.x % vpr;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > npix) p1 = npix;
    const float gs = *gscale;
    float acc[NC][VEC];
    float accb[NC];
    float wv[NC][VEC];
#pragma unroll
    for (int o = 0; o < NC; ++o) {
        accb[o] = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) { acc[o][i] = 0.f; wv[o][i] = (o < nc && r < rows) ? w[o * C + cv * VEC + i] : 0.f; }
    }
    if (r < rows) {
        // (image, pixel-in-image) of p, advanced incrementally: a 64-bit divide per pixel costs more than the FMAs
        long b = (p0 + r) / HW, hw = (p0 + r) % HW;
#pragma unroll 2
        for (long p = p0 + r; p < p1; p += rows) {
            float gv[NC];
#pragma unroll
            for (int o = 0; o < NC; ++o) gv[o] = o < nc ? g[(b * nc + o) * HW + hw] : 0.f;
            hw += rows;
            while (hw >= HW) { hw -= HW; ++b; }
            float v[VEC], d[VEC];
            load_vec(x + p * ldx + cv * VEC, v);
#pragma unroll
            for (int i = 0; i < VEC; ++i) d[i] = 0.f;
#pragma unroll
            for (int o = 0; o < NC; ++o) {
                if (cv == 0) accb[o] += gv[o];
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    d[i] = fmaf(gv[o], wv[o][i], d[i]);
                    acc[o][i] = fmaf(gv[o], v[i], acc[o][i]);
                }
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) d[i] *= gs;
            store_vec(dx + p * lddx + cv * VEC, d);
        }
    }
    float* out = parts + (long)blockIdx.x * (NC * C + NC);
    block_reduce_rows<NC, VEC>(acc, C, vpr, rows, out);
    // bias partial: only cv == 0 threads hold it
    __shared__ float bred[kThreads];
#pragma unroll
    for (int o = 0; o < NC; ++o) {
        __syncthreads();
        bred[threadIdx.x] = (r < rows && cv == 0) ? accb[o] : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < kThreads; ++i) t += bred[i];
            out[NC * C + o] = t;
        }
    }
}
// outc forward / backward as shared-memory streams (bf16, dense, C = 64, at most 2 classes, H*W a multiple of 128):
// 128-pixel tiles of x by cp.async.bulk; 8 threads share a pixel (16 B each) as in the kernels above.  Forward stages
// the 2 x 128 logits of a tile and writes them as two 512-byte runs; backward also fetches the tile's dlogits rows by
// bulk copy, overwrites the x tile with dx and stores it in bulk.  Same arithmetic as the register kernels.
constexpr int kOcTile = 16384, kOcPix = 128;
constexpr int kOcFwdStages = 4;
constexpr int kOcFwdSmem = kOcFwdStages * kOcTile;
__global__ void __launch_bounds__(kThreads, 3) outc_fwd_stream_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, int nc,
                                                                      float* __restrict__ logits, long HW, long chunk) {
    extern __shared__ __align__(128) uint8_t st_smem[];
    __shared__ __align__(8) uint64_t full_bar[kOcFwdStages];
    __shared__ float res[2][kOcPix];
    constexpr int C = 64, VEC = 8;
    const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > HW) p1 = HW;
    const int b = blockIdx.y;
    const long nbytes = (p1 - p0) * C * 2;
    const int ntiles = (int)((nbytes + kOcTile - 1) / kOcTile);
    const uint8_t* xb = reinterpret_cast<const uint8_t*>(x) + ((long)b * HW + p0) * C * 2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kOcFwdStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const long o = (long)t * kOcTile;
        const uint32_t bytes = (uint32_t)(nbytes - o < kOcTile ? nbytes - o : kOcTile);
        mbar_expect_tx(&full_bar[t % kOcFwdStages], bytes);
        bulk_load_1d(st_smem + (t % kOcFwdStages) * kOcTile, xb + o, bytes, &full_bar[t % kOcFwdStages]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kOcFwdStages && t < ntiles; ++t) issue(t);
    float wr[2][VEC];
#pragma unroll
    for (int o = 0; o < 2; ++o)
#pragma unroll
        for (int i = 0; i < VEC; ++i) wr[o][i] = o < nc ? w[o * C + sub * VEC + i] : 0.f;
    const float bo = (threadIdx.x >> 7) < nc ? bias[threadIdx.x >> 7] : 0.f;
    for (int t = 0; t < ntiles; ++t) {
        const long o = (long)t * kOcTile;
        const int npx = (int)((nbytes - o < kOcTile ? nbytes - o : kOcTile) / (C * 2));
        mbar_wait(&full_bar[t % kOcFwdStages], (uint32_t)((t / kOcFwdStages) & 1));
        const bf16* xs = reinterpret_cast<const bf16*>(st_smem + (t % kOcFwdStages) * kOcTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int px = pl + 32 * k;
            float acc0 = 0.f, acc1 = 0.f;
            if (px < npx) {
                float v[VEC];
                load_vec(xs + px * C + sub * VEC, v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) { acc0 = fmaf(v[i], wr[0][i], acc0); acc1 = fmaf(v[i], wr[1][i], acc1); }
            }
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 4); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 4);
            if (sub == 0) { res[0][px] = acc0; res[1][px] = acc1; }
        }
        __syncthreads();                                       // tile consumed, results staged
        if (threadIdx.x == 0 && t + kOcFwdStages < ntiles) issue(t + kOcFwdStages);
        {
            const int o2 = threadIdx.x >> 7, px = threadIdx.x & 127;
            if (o2 < nc && px < npx) logits[((long)b * nc + o2) * HW + p0 + (long)t * kOcPix + px] = res[o2][px] + bo;
        }
        __syncthreads();                                       // res may be overwritten
    }
}

constexpr int kOcBwdStages = 4;
constexpr int kOcBwdStage = kOcTile + 2 * kOcPix * 4;          // x tile | two dlogits rows
constexpr int kOcBwdSmem = kOcBwdStages * kOcBwdStage;
__global__ void __launch_bounds__(kThreads, 3) outc_bwd_stream_kernel(const float* __restrict__ g, const float* __restrict__ gscale,
                                                                      const bf16* __restrict__ x, bf16* __restrict__ dx,
                                                                      const float* __restrict__ w, int nc, long HW, long chunk,
                                                                      float* __restrict__ parts) {
    extern __shared__ __align__(128) uint8_t st_smem[];
    __shared__ __align__(8) uint64_t full_bar[kOcBwdStages];
    constexpr int C = 64, VEC = 8, NC = 2;
    const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > HW) p1 = HW;
    const int b = blockIdx.y;
    const long nbytes = (p1 - p0) * C * 2;
    const int ntiles = (int)((nbytes + kOcTile - 1) / kOcTile);
    const uint8_t* xb = reinterpret_cast<const uint8_t*>(x) + ((long)b * HW + p0) * C * 2;
    uint8_t* ob = reinterpret_cast<uint8_t*>(dx) + ((long)b * HW + p0) * C * 2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kOcBwdStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const long o = (long)t * kOcTile;
        const uint32_t bytes = (uint32_t)(nbytes - o < kOcTile ? nbytes - o : kOcTile);
        const uint32_t gb = bytes / (C * 2) * 4;                         // dlogits bytes per class (a multiple of 16)
        uint8_t* st = st_smem + (t % kOcBwdStages) * kOcBwdStage;
        mbar_expect_tx(&full_bar[t % kOcBwdStages], bytes + nc * gb);
        bulk_load_1d(st, xb + o, bytes, &full_bar[t % kOcBwdStages]);
        for (int o2 = 0; o2 < nc; ++o2)
            bulk_load_1d(st + kOcTile + o2 * kOcPix * 4, g + ((long)b * nc + o2) * HW + p0 + (long)t * kOcPix, gb,
                         &full_bar[t % kOcBwdStages]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kOcBwdStages - 1 && t < ntiles; ++t) issue(t);
    const float gs = *gscale;
    float acc[NC][VEC], accb[NC], wv[NC][VEC];
#pragma unroll
    for (int o = 0; o < NC; ++o) {
        accb[o] = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) { acc[o][i] = 0.f; wv[o][i] = o < nc ? w[o * C + sub * VEC + i] : 0.f; }
    }
    for (int t = 0; t < ntiles; ++t) {
        const long o = (long)t * kOcTile;
        const int bytes = (int)(nbytes - o < kOcTile ? nbytes - o : kOcTile);
        const int npx = bytes / (C * 2);
        uint8_t* st = st_smem + (t % kOcBwdStages) * kOcBwdStage;
        mbar_wait(&full_bar[t % kOcBwdStages], (uint32_t)((t / kOcBwdStages) & 1));
        bf16* xs = reinterpret_cast<bf16*>(st);
        const float* gsm = reinterpret_cast<const float*>(st + kOcTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int px = pl + 32 * k;
            if (px < npx) {
                float gv[NC];
#pragma unroll
                for (int o2 = 0; o2 < NC; ++o2) gv[o2] = o2 < nc ? gsm[o2 * kOcPix + px] : 0.f;
                float v[VEC], d[VEC];
                load_vec(xs + px * C + sub * VEC, v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) d[i] = 0.f;
#pragma unroll
                for (int o2 = 0; o2 < NC; ++o2) {
                    if (sub == 0) accb[o2] += gv[o2];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        d[i] = fmaf(gv[o2], wv[o2][i], d[i]);
                        acc[o2][i] = fmaf(gv[o2], v[i], acc[o2][i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < VEC; ++i) d[i] *= gs;
                store_vec(xs + px * C + sub * VEC, d);
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            bulk_store_1d(ob + o, xs, (uint32_t)bytes);
            tma_store_commit();
            if (t >= 1) {
                tma_store_wait_read1();
                if (t - 1 + kOcBwdStages < ntiles) issue(t - 1 + kOcBwdStages);
            } else if (kOcBwdStages - 1 < ntiles) {
                issue(kOcBwdStages - 1);
            }
        }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
    float* out = parts + ((long)blockIdx.y * gridDim.x + blockIdx.x) * (NC * C + NC);
    block_reduce_rows<NC, VEC>(acc, C, 8, kThreads / 8, out);
    __shared__ float bred[kThreads];
#pragma unroll
    for (int o2 = 0; o2 < NC; ++o2) {
        __syncthreads();
        bred[threadIdx.x] = sub == 0 ? accb[o2] : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tsum = 0.f;
            for (int i = 0; i < kThreads; ++i) tsum += bred[i];
            out[NC * C + o2] = tsum;
        }
    }
}
// finalize: dW[o][c] (nc x C) and db[o] from parts laid out with the padded NC
// ---------------------------------------------------------------------------------------------------------
// Output head fused with the last DoubleConv's elementwise passes (bf16, C = 64, nc <= 2, HW % 128 == 0).
// The 1x1 outc conv (UCA:125,162) reads and its backward writes a [B,H,W,64] tensor at full resolution: the block output
// h = relu(bn2(y2)) * s going in (written by se_scale, read by outc_fwd, read again by outc_bwd), and its gradient
// dh = W^T dlogits coming back (written by outc_bwd, read by the SE+BN2 reduction and by the BN2 apply pass) — 2.1 GB
// each at batch 64, five passes, none of which holds information: h is a per-pixel function of y2, dh of the 8 bytes of
// dlogits per pixel.  So neither tensor exists any more:
//   forward   se_scale_outc_fwd:      y2 -> logits                       (h formed per pixel, rounded to bf16 as before)
//   backward  outc_bn_bwd_reduce:     (dlogits, y2) -> S1, S2 partial sums of the SE+BN2 backward, and dW_outc, db_outc
//             outc_bn_bwd_apply:      (dlogits, y2) -> dY2
// h and dh stay in fp32 registers (the unfused passes rounded them to bf16 on their way through memory), so the fused head
// is slightly CLOSER to the fp32 reference than the passes it replaces.  The kernels are sized for the ALU, not only for
// HBM: at 64 channels a pass has ~9 instructions per element before it becomes issue-bound, so per-(image, channel) factors
// are folded into the weights / taken out of the sums (s into W_outc, gscale into W_outc^T, the mean out of sum dz*(y-mean)).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 3) se_scale_outc_fwd_stream_kernel(const bf16* __restrict__ y, const float* __restrict__ scale,
                                                                               const float* __restrict__ shift, const float* __restrict__ s,
                                                                               const float* __restrict__ w, const float* __restrict__ bias,
                                                                               int nc, float* __restrict__ logits, long HW, long chunk) {
    extern __shared__ __align__(128) uint8_t st_smem[];
    __shared__ __align__(8) uint64_t full_bar[kOcFwdStages];
    __shared__ float res[2][kOcPix];
    constexpr int C = 64, VEC = 8;
    const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > HW) p1 = HW;
    const int b = blockIdx.y;
    const long nbytes = (p1 - p0) * C * 2;
    const int ntiles = (int)((nbytes + kOcTile - 1) / kOcTile);
    const uint8_t* xb = reinterpret_cast<const uint8_t*>(y) + ((long)b * HW + p0) * C * 2;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kOcFwdStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const long o = (long)t * kOcTile;
        const uint32_t bytes = (uint32_t)(nbytes - o < kOcTile ? nbytes - o : kOcTile);
        mbar_expect_tx(&full_bar[t % kOcFwdStages], bytes);
        bulk_load_1d(st_smem + (t % kOcFwdStages) * kOcTile, xb + o, bytes, &full_bar[t % kOcFwdStages]);
    };
    if (threadIdx.x == 0)
        for (int t = 0; t < kOcFwdStages && t < ntiles; ++t) issue(t);
    float wr[2][VEC], a[VEC], sh[VEC], gg[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        a[i] = scale[sub * VEC + i]; sh[i] = shift[sub * VEC + i];
        gg[i] = s ? s[(long)b * C + sub * VEC + i] : 1.f;
#pragma unroll
        for (int o = 0; o < 2; ++o) wr[o][i] = o < nc ? w[o * C + sub * VEC + i] * gg[i] : 0.f;      // W_outc * s[b, c]
    }
    const float bo = (threadIdx.x >> 7) < nc ? bias[threadIdx.x >> 7] : 0.f;
    for (int t = 0; t < ntiles; ++t) {
        const long o = (long)t * kOcTile;
        const int npx = (int)((nbytes - o < kOcTile ? nbytes - o : kOcTile) / (C * 2));
        mbar_wait(&full_bar[t % kOcFwdStages], (uint32_t)((t / kOcFwdStages) & 1));
        const bf16* xs = reinterpret_cast<const bf16*>(st_smem + (t % kOcFwdStages) * kOcTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int px = pl + 32 * k;
            float acc0 = 0.f, acc1 = 0.f;
            if (px < npx) {
                float v[VEC];
                load_vec(xs + px * C + sub * VEC, v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) {
                    const float h = fmaxf(fmaf(a[i], v[i], sh[i]), 0.f);
                    acc0 = fmaf(h, wr[0][i], acc0); acc1 = fmaf(h, wr[1][i], acc1);
                }
            }
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
            acc0 += __shfl_xor_sync(0xffffffffu, acc0, 4); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 4);
            if (sub == 0) { res[0][px] = acc0; res[1][px] = acc1; }
        }
        __syncthreads();                                       // tile consumed, results staged
        if (threadIdx.x == 0 && t + kOcFwdStages < ntiles) issue(t + kOcFwdStages);
        {
            const int o2 = threadIdx.x >> 7, px = threadIdx.x & 127;
            if (o2 < nc && px < npx) logits[((long)b * nc + o2) * HW + p0 + (long)t * kOcPix + px] = res[o2][px] + bo;
        }
        __syncthreads();                                       // res may be overwritten
    }
}

// MODE 0: reduction (no store): parts_bn [(b*gridDim.x + blk)][2][C] = (sum m*dh, sum m*dh*(y - mean)), m = (a*y + b > 0);
//         parts_oc [(b*gridDim.x + blk)][2*C + 2] = (sum_p g_o * h, sum_p g_o) for the outc weight / bias gradients.
// MODE 1: apply: dY2 = g*dz - y*k2 + k0 as in bn_bwd_apply_stream_kernel, dz = m * (dh*(g*s) + g*dp/HW) (SE) or m * dh * g.
template <int MODE, bool SE>
__global__ void __launch_bounds__(kThreads, MODE == 0 ? 2 : 3) outc_bn_bwd_stream_kernel(const float* __restrict__ g, const float* __restrict__ gscale,
                                                                         const float* __restrict__ w, int nc,
                                                                         const bf16* __restrict__ y, bf16* __restrict__ dy, long HW,
                                                                         long chunk, float inv_hw, const float* __restrict__ scale,
                                                                         const float* __restrict__ shift, const float* __restrict__ mean,
                                                                         const float* __restrict__ invstd, const float* __restrict__ s,
                                                                         const float* __restrict__ dp, const float* __restrict__ coef,
                                                                         float* __restrict__ parts_bn, float* __restrict__ parts_oc) {
    extern __shared__ __align__(128) uint8_t st_smem[];
    __shared__ __align__(8) uint64_t full_bar[kOcBwdStages];
    constexpr int C = 64, VEC = 8, NC = 2;
    const int sub = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const long p0 = (long)blockIdx.x * chunk;
    long p1 = p0 + chunk; if (p1 > HW) p1 = HW;
    const int b = blockIdx.y;
    const long nbytes = p1 > p0 ? (p1 - p0) * C * 2 : 0;
    const int ntiles = (int)((nbytes + kOcTile - 1) / kOcTile);
    const uint8_t* xb = reinterpret_cast<const uint8_t*>(y) + ((long)b * HW + p0) * C * 2;
    uint8_t* ob = MODE == 1 ? reinterpret_cast<uint8_t*>(dy) + ((long)b * HW + p0) * C * 2 : nullptr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kOcBwdStages; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int t) {
        const long o = (long)t * kOcTile;
        const uint32_t bytes = (uint32_t)(nbytes - o < kOcTile ? nbytes - o : kOcTile);
        const uint32_t gb = bytes / (C * 2) * 4;                         // dlogits bytes per class (a multiple of 16)
        uint8_t* st = st_smem + (t % kOcBwdStages) * kOcBwdStage;
        mbar_expect_tx(&full_bar[t % kOcBwdStages], bytes + nc * gb);
        bulk_load_1d(st, xb + o, bytes, &full_bar[t % kOcBwdStages]);
        for (int o2 = 0; o2 < nc; ++o2)
            bulk_load_1d(st + kOcTile + o2 * kOcPix * 4, g + ((long)b * nc + o2) * HW + p0 + (long)t * kOcPix, gb,
                         &full_bar[t % kOcBwdStages]);
    };
    // MODE 0 refills a stage as soon as every thread has consumed it (all stages in flight); MODE 1 after its store has read it
    if (threadIdx.x == 0)
        for (int t = 0; t < (MODE == 0 ? kOcBwdStages : kOcBwdStages - 1) && t < ntiles; ++t) issue(t);
    const float gs = *gscale;
    float wv[NC][VEC], a[VEC], sh[VEC], gg[VEC];
    float acc[NC][VEC], accb[NC], sbn[2][VEC], mu[VEC], m0[VEC], m1[VEC], k2[VEC], k0[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = sub * VEC + i;
        a[i] = scale[c]; sh[i] = shift[c];
        gg[i] = s ? s[(long)b * C + c] : 1.f;
#pragma unroll
        for (int o = 0; o < NC; ++o) { wv[o][i] = o < nc ? w[o * C + c] * gs : 0.f; acc[o][i] = 0.f; }      // W_outc^T * gscale
        sbn[0][i] = sbn[1][i] = 0.f;
        mu[i] = MODE == 0 ? mean[c] : 0.f;
        if (MODE == 1) {
            const float gc = coef[c], c1 = coef[C + c], c2 = coef[2 * C + c];
            const float sv = SE ? s[(long)b * C + c] : 1.f, dpv = SE ? dp[(long)b * C + c] * inv_hw : 0.f;
            k2[i] = gc * c2 * invstd[c];
            k0[i] = mean[c] * k2[i] - gc * c1;
            m0[i] = gc * sv; m1[i] = gc * dpv;
        } else { k2[i] = k0[i] = m0[i] = m1[i] = 0.f; }
    }
    accb[0] = accb[1] = 0.f;
    for (int t = 0; t < ntiles; ++t) {
        const long o = (long)t * kOcTile;
        const int bytes = (int)(nbytes - o < kOcTile ? nbytes - o : kOcTile);
        const int npx = bytes / (C * 2);
        uint8_t* st = st_smem + (t % kOcBwdStages) * kOcBwdStage;
        mbar_wait(&full_bar[t % kOcBwdStages], (uint32_t)((t / kOcBwdStages) & 1));
        bf16* xs = reinterpret_cast<bf16*>(st);
        const float* gsm = reinterpret_cast<const float*>(st + kOcTile);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int px = pl + 32 * k;
            if (px < npx) {
                float gv[NC];
#pragma unroll
                for (int o2 = 0; o2 < NC; ++o2) gv[o2] = o2 < nc ? gsm[o2 * kOcPix + px] : 0.f;
                float v[VEC], d[VEC];
                load_vec(xs + px * C + sub * VEC, v);
#pragma unroll
                for (int i = 0; i < VEC; ++i) d[i] = fmaf(gv[1], wv[1][i], gv[0] * wv[0][i]);          // dh = W^T g * gscale
                if (MODE == 0) {
#pragma unroll
                    for (int o2 = 0; o2 < NC; ++o2) if (sub == 0) accb[o2] += gv[o2];
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float pre = fmaf(a[i], v[i], sh[i]);
                        if (pre > 0.f) {
                            // relu(pre) = pre here; the SE scale s[b,c] multiplies the finished sums (one image per block)
                            acc[0][i] = fmaf(gv[0], pre, acc[0][i]);
                            acc[1][i] = fmaf(gv[1], pre, acc[1][i]);
                            sbn[0][i] += d[i];
                            sbn[1][i] = fmaf(d[i], v[i], sbn[1][i]);                                   // the mean is taken out at the end
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < VEC; ++i) {
                        const float da = SE ? fmaf(d[i], m0[i], m1[i]) : d[i] * m0[i];
                        const float dz = fmaf(a[i], v[i], sh[i]) > 0.f ? da : 0.f;
                        d[i] = fmaf(-v[i], k2[i], dz) + k0[i];
                    }
                    store_vec(xs + px * C + sub * VEC, d);
                }
            }
        }
        if (MODE == 1) {
            fence_proxy_async_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                bulk_store_1d(ob + o, xs, (uint32_t)bytes);
                tma_store_commit();
                if (t >= 1) {
                    tma_store_wait_read1();
                    if (t - 1 + kOcBwdStages < ntiles) issue(t - 1 + kOcBwdStages);
                } else if (kOcBwdStages - 1 < ntiles) {
                    issue(kOcBwdStages - 1);
                }
            }
        } else {
            __syncthreads();                                   // every thread is done with this stage
            if (threadIdx.x == 0 && t + kOcBwdStages < ntiles) issue(t + kOcBwdStages);
        }
    }
    if (MODE == 1) {
        if (threadIdx.x == 0) tma_store_wait_all();
        return;
    }
    const long row = (long)blockIdx.y * gridDim.x + blockIdx.x;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        sbn[1][i] = fmaf(-mu[i], sbn[0][i], sbn[1][i]);          // sum dz*(y - mean) = sum dz*y - mean * sum dz
        acc[0][i] *= gg[i]; acc[1][i] *= gg[i];                  // h = relu(pre) * s[b,c]
    }
    block_reduce_rows<2, VEC>(sbn, C, 8, kThreads / 8, parts_bn + row * 2 * C);
    float* out = parts_oc + row * (NC * C + NC);
    block_reduce_rows<NC, VEC>(acc, C, 8, kThreads / 8, out);
    __shared__ float bred[kThreads];
#pragma unroll
    for (int o2 = 0; o2 < NC; ++o2) {
        __syncthreads();
        bred[threadIdx.x] = sub == 0 ? accb[o2] : 0.f;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tsum = 0.f;
            for (int i = 0; i < kThreads; ++i) tsum += bred[i];
            out[NC * C + o2] = tsum;
        }
    }
}

__global__ void outc_bwd_finalize_kernel(const float* __restrict__ parts, int nparts, int NCpad, int nc, int C,
                                         const float* __restrict__ gscale, float* __restrict__ dw,
                                         float* __restrict__ db) {
    // one warp per output element; lanes stride over the partial rows
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= nc * C + nc) return;
    const int stride = NCpad * C + NCpad;
    const int src = i < nc * C ? i : NCpad * C + (i - nc * C);
    double t = 0.0;
    for (int p = lane; p < nparts; p += 32) t += (double)parts[(long)p * stride + src];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
        const float v = (float)(t * (double)*gscale);
        if (i < nc * C) dw[i] = v; else db[i - nc * C] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// softmax cross-entropy, ignore_index, mean over valid pixels.                         UCA:465,344
//   g[b,o,h,w] = (softmax - onehot) * valid   (un-normalised dlogits, fp32 NCHW)
//   parts[blk] = {sum of -log p[target] over valid pixels, number of valid pixels}
// ce_finalize: loss = sum/N (0/0 -> NaN like torch), gscale = upstream/N.
// Also emits the argmax class map (first maximum wins, UCA:220) when `mask` is non-null.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) ce_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                      int nc, long npix, long HW, long long ignore_index,
                                                      float* __restrict__ g, long long* __restrict__ mask,
                                                      float* __restrict__ parts) {
    float lsum = 0.f, lcnt = 0.f;
    for (long p = (long)blockIdx.x * kThreads + threadIdx.x; p < npix; p += (long)gridDim.x * kThreads) {
        const long b = p / HW, hw = p % HW;
        const float* lp = logits + b * nc * HW + hw;
        float m = lp[0];
        int am = 0;
        for (int o = 1; o < nc; ++o) {
            const float v = lp[o * HW];
            if (v > m || (v != v && m == m)) { m = v; am = o; }   // first max wins; NaN propagates like torch.max
        }
        if (mask) mask[p] = am;
        if (!target) continue;
        float se = 0.f;
        for (int o = 0; o < nc; ++o) se += expf(lp[o * HW] - m);
        const float lse = m + logf(se);
        long long t = target[p];
        const bool valid = t != ignore_index;
        if (valid && (t < 0 || t >= nc)) { lsum = nanf(""); t = 0; }   // torch asserts here; poison the loss instead
        if (valid) { lsum += lse - lp[t * HW]; lcnt += 1.f; }
        if (g) {
            float* gp = g + b * nc * HW + hw;
            for (int o = 0; o < nc; ++o) {
                const float pr = expf(lp[o * HW] - lse);
                gp[o * HW] = valid ? pr - (o == t ? 1.f : 0.f) : 0.f;
            }
        }
    }
    __shared__ float rs[kThreads / 32], rc[kThreads / 32];
    lsum = warp_sum(lsum);
    lcnt = warp_sum(lcnt);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = lsum; rc[threadIdx.x >> 5] = lcnt; }
    __syncthreads();
    if (threadIdx.x == 0 && parts) {
        float a = 0.f, c = 0.f;
        for (int i = 0; i < kThreads / 32; ++i) { a += rs[i]; c += rc[i]; }
        parts[2 * blockIdx.x] = a;
        parts[2 * blockIdx.x + 1] = c;
    }
}
// Two classes, HW % 4 == 0: the same arithmetic as ce_kernel, in the same order per pixel, on four consecutive pixels per
// thread with 16-byte loads / stores (two float4 of logits, two longlong2 of labels in; two float4 of dlogits out).
__global__ void __launch_bounds__(kThreads) ce2_vec_kernel(const float* __restrict__ logits, const long long* __restrict__ target,
                                                           long nquad, long HW, long long ignore_index, float* __restrict__ g,
                                                           long long* __restrict__ mask, float* __restrict__ parts) {
    float lsum = 0.f, lcnt = 0.f;
    for (long q = (long)blockIdx.x * kThreads + threadIdx.x; q < nquad; q += (long)gridDim.x * kThreads) {
        const long p = q * 4;
        const long b = p / HW, hw = p % HW;
        const float* lp = logits + b * 2 * HW + hw;
        const float4 l0 = *reinterpret_cast<const float4*>(lp), l1 = *reinterpret_cast<const float4*>(lp + HW);
        const float a[4] = {l0.x, l0.y, l0.z, l0.w}, c[4] = {l1.x, l1.y, l1.z, l1.w};
        long long t[4] = {0, 0, 0, 0};
        if (target) {
            const longlong2 t01 = *reinterpret_cast<const longlong2*>(target + p), t23 = *reinterpret_cast<const longlong2*>(target + p + 2);
            t[0] = t01.x; t[1] = t01.y; t[2] = t23.x; t[3] = t23.y;
        }
        float g0[4], g1[4];
        long long am[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float m = a[i];
            am[i] = 0;
            if (c[i] > m || (c[i] != c[i] && m == m)) { m = c[i]; am[i] = 1; }     // first max wins; NaN propagates like torch.max
            float se = 0.f;
            se += expf(a[i] - m);
            se += expf(c[i] - m);
            const float lse = m + logf(se);
            long long ti = t[i];
            const bool valid = ti != ignore_index;
            if (target && valid && (ti < 0 || ti >= 2)) { lsum = nanf(""); ti = 0; }
            if (target && valid) { lsum += lse - (ti ? c[i] : a[i]); lcnt += 1.f; }
            g0[i] = valid ? expf(a[i] - lse) - (ti == 0 ? 1.f : 0.f) : 0.f;
            g1[i] = valid ? expf(c[i] - lse) - (ti == 1 ? 1.f : 0.f) : 0.f;
        }
        if (mask) {
            *reinterpret_cast<longlong2*>(mask + p) = make_longlong2(am[0], am[1]);
            *reinterpret_cast<longlong2*>(mask + p + 2) = make_longlong2(am[2], am[3]);
        }
        if (target && g) {
            float* gp = g + b * 2 * HW + hw;
            *reinterpret_cast<float4*>(gp) = make_float4(g0[0], g0[1], g0[2], g0[3]);
            *reinterpret_cast<float4*>(gp + HW) = make_float4(g1[0], g1[1], g1[2], g1[3]);
        }
    }
    __shared__ float rs[kThreads / 32], rc[kThreads / 32];
    lsum = warp_sum(lsum);
    lcnt = warp_sum(lcnt);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = lsum; rc[threadIdx.x >> 5] = lcnt; }
    __syncthreads();
    if (threadIdx.x == 0 && parts) {
        float sa = 0.f, sc = 0.f;
        for (int i = 0; i < kThreads / 32; ++i) { sa += rs[i]; sc += rc[i]; }
        parts[2 * blockIdx.x] = sa;
        parts[2 * blockIdx.x + 1] = sc;
    }
}

__global__ void ce_finalize_kernel(const float* __restrict__ parts, int nparts, const float* __restrict__ upstream,
                                   float* __restrict__ loss, float* __restrict__ gscale) {
    // one warp
    const int lane = threadIdx.x;
    double a = 0.0, c = 0.0;
    for (int i = lane; i < nparts; i += 32) { a += (double)parts[2 * i]; c += (double)parts[2 * i + 1]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
    if (lane != 0) return;
    const float n = (float)c;
    loss[0] = (float)a / n;
    loss[1] = n;
    gscale[0] = (upstream ? *upstream : 1.f) / n;
}

// ---------------------------------------------------------------------------------------------------------
// layout conversion at the module boundary and weight packing
// ---------------------------------------------------------------------------------------------------------
// compute_metrics (UCA:214-269) on the device: argmax of the class logits (first maximum wins, like
// torch.max(outputs, 1) at UCA:220) against the label map, pixels with label == ignore_index dropped (UCA:223),
// counted into a (nc+1) x nc table: row = label (row nc = any other label value), column = predicted class.
// TP / FP / FN of the reference are sums over that table, so only (nc+1)*nc integers travel to the host instead of
// the two masks (UCA:229-230).  Two-stage and deterministic: per-block tables, then one summing block.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxMetricClasses = 8;
__global__ void __launch_bounds__(kThreads) confusion_kernel(const float* __restrict__ logits,
                                                             const long long* __restrict__ target, int nc, long npix,
                                                             long HW, long long ignore_index,
                                                             unsigned long long* __restrict__ parts) {
    __shared__ unsigned int cnt[(kMaxMetricClasses + 1) * kMaxMetricClasses];
    const int ncell = (nc + 1) * nc;
    for (int i = threadIdx.x; i < ncell; i += kThreads) cnt[i] = 0u;
    __syncthreads();
    for (long p = (long)blockIdx.x * kThreads + threadIdx.x; p < npix; p += (long)gridDim.x * kThreads) {
        const long long t = target[p];
        if (t == ignore_index) continue;
        const long b = p / HW, hw = p % HW;
        const float* lp = logits + b * nc * HW + hw;
        float m = lp[0];
        int am = 0;
        for (int o = 1; o < nc; ++o) {
            const float v = lp[o * HW];
            if (v > m || (v != v && m == m)) { m = v; am = o; }
        }
        const int row = (t >= 0 && t < nc) ? (int)t : nc;
        atomicAdd(&cnt[row * nc + am], 1u);             // shared-memory integer counter: order-independent
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ncell; i += kThreads) parts[(long)blockIdx.x * ncell + i] = cnt[i];
}
__global__ void confusion_finalize_kernel(const unsigned long long* __restrict__ parts, int nparts, int ncell,
                                          long long* __restrict__ counts) {
    const int i = threadIdx.x;
    if (i >= ncell) return;
    unsigned long long t = 0;
    for (int k = 0; k < nparts; ++k) t += parts[(long)k * ncell + i];
    counts[i] = (long long)t;
}

// ---------------------------------------------------------------------------------------------------------
// Device side of the reference's input preprocessing (UCA:200-210, 428-433): uint8 tile -> T.ToTensor() (/255) ->
// T.Normalize([mean],[std]); uint8 mask -> T.ToTensor()(mask).long(), i.e. 255 -> 1 and everything else -> 0.
// Same IEEE operations in the same order as torchvision, so the tensors are bit-identical; 16 pixels per thread.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) prep_u8_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ mask,
                                                           float* __restrict__ out, long long* __restrict__ lab, long n,
                                                           float mean, float stdv) {
    const long i0 = ((long)blockIdx.x * kThreads + threadIdx.x) * 16;
    if (i0 >= n) return;
    if (i0 + 16 <= n && ((reinterpret_cast<uintptr_t>(img + i0) | reinterpret_cast<uintptr_t>(mask + i0)) & 15) == 0) {
        const uint4 a = *reinterpret_cast<const uint4*>(img + i0);
        const uint4 m = *reinterpret_cast<const uint4*>(mask + i0);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float v = (float)((aw[k >> 2] >> (8 * (k & 3))) & 0xff) / 255.f;
            out[i0 + k] = (v - mean) / stdv;
            lab[i0 + k] = (long long)((float)((mw[k >> 2] >> (8 * (k & 3))) & 0xff) / 255.f);
        }
    } else {
        for (long i = i0; i < n && i < i0 + 16; ++i) {
            out[i] = ((float)img[i] / 255.f - mean) / stdv;
            lab[i] = (long long)((float)mask[i] / 255.f);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// im2col of the (B,Cin,H,W) NCHW fp32 network input for the first 3x3 conv: col[p][tap*Cin + c], zero padded to
// Kpad columns (K = 9*Cin is not a multiple of the MMA K; the padding exists only in this staging buffer).
template <typename T, int CIN>
__global__ void __launch_bounds__(256) im2col3x3_small_kernel(const float* __restrict__ x, T* __restrict__ col, int B, int H,
                                                               int W, int Kpad) {
    // Cin <= 4: one thread per pixel gathers its 9*Cin inputs with warp-coalesced loads (lanes = consecutive w),
    // then writes the whole Kpad-wide row as 16-byte stores.
    constexpr int VEC = VecTraits<T>::N;
    constexpr int K = 9 * CIN;
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long)B * H * W) return;
    const int w = (int)(p % W), h = (int)((p / W) % H);
    const long b = p / ((long)W * H);
    float v[K];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
        const bool in = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
        for (int c = 0; c < CIN; ++c) v[tap * CIN + c] = in ? __ldg(x + ((b * CIN + c) * H + hh) * W + ww) : 0.f;
    }
    T* dst = col + p * Kpad;
#pragma unroll
    for (int j = 0; j < (K + VEC - 1) / VEC; ++j) {
        float o[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] = (j * VEC + e) < K ? v[(j * VEC + e) < K ? (j * VEC + e) : 0] : 0.f;
        store_vec(dst + j * VEC, o);
    }
    float z[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) z[e] = 0.f;
    for (int j = (K + VEC - 1) / VEC; j < Kpad / VEC; ++j) store_vec(dst + j * VEC, z);
}
// ---------------------------------------------------------------------------------------------------------
// First conv (Cin <= 5) in the row-pair layout of the tcgen05 path: instead of one im2col row of 9*Cin values per
// pixel, one row per PIXEL PAIR (rows 2i, 2i+1 of column x) holding the 4 x 3 patch they share:
//   colp[(b,i,x)][(vr*3 + kw)*Cin + c] = x[b, c, 2i + vr - 1, x + kw - 1]     (zero outside the image / beyond 12*Cin)
// With the pair-packed filter  wp[(j,o)][(vr*3+kw)*Cin + c] = W[o][c][vr - j][kw]  (zero when vr - j is not 0..2) the
// layer is ONE GEMM  out[(2i+j, x)][o] = sum_k wp[(j,o)][k] * colp[(b,i,x)][k]  with M = 128, N = 256 pixel pairs:
// half the staging traffic of the per-pixel im2col and full-rate 128x256x16 MMAs.                      UCA:81
// ---------------------------------------------------------------------------------------------------------
template <typename T, int CIN>
__global__ void __launch_bounds__(256) im2col_pairs_kernel(const float* __restrict__ x, T* __restrict__ colp, int B, int H,
                                                           int W) {
    constexpr int VEC = VecTraits<T>::N;
    constexpr int K = 12 * CIN;
    const int HP = H >> 1;
    const long n = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = n < (long)B * HP * W;                 // (every thread reaches the barrier below)
    const int xw = (int)(n % W), i = (int)((n / W) % HP);
    const long b = n / ((long)W * HP);
    float v[K];
#pragma unroll
    for (int vr = 0; vr < 4; ++vr) {
        const int hh = 2 * i + vr - 1;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int ww = xw + kw - 1;
            const bool in = live && hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
            for (int c = 0; c < CIN; ++c)
                v[(vr * 3 + kw) * CIN + c] = in ? __ldg(x + ((b * CIN + c) * H + hh) * W + ww) : 0.f;
        }
    }
    if constexpr (sizeof(T) == 4) {
        // fp32 rows (cross-check only): each thread writes its own 256-byte row
        if (!live) return;
        T* dst = colp + n * 64;
#pragma unroll
        for (int j = 0; j < 64 / VEC; ++j) {
            float o[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) o[e] = (j * VEC + e) < K ? v[(j * VEC + e) < K ? (j * VEC + e) : 0] : 0.f;
            store_vec(dst + j * VEC, o);
        }
    } else {
    // rows are staged in shared memory and leave as contiguous 16-byte chunks (a warp writes 512 contiguous bytes): a
    // thread writing its own row would touch 32 different 128-byte lines per store instruction
    __shared__ __align__(16) T stage[256 * 64];
    // chunk j of row r at slot (j ^ (r & 7)): conflict-free both for the row-wise writes and the chunk-wise reads (bf16)
    constexpr int NCH = 64 / VEC;
    T* srow = stage + threadIdx.x * 64;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        float o[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) o[e] = (j * VEC + e) < K ? v[(j * VEC + e) < K ? (j * VEC + e) : 0] : 0.f;
        store_vec(srow + ((j ^ (threadIdx.x & (NCH - 1))) * VEC), o);
    }
    __syncthreads();
    const long row0 = (long)blockIdx.x * blockDim.x;
    const long nrows = (long)B * HP * W;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int cidx = k * 256 + threadIdx.x;              // chunk index within the block's [256][NCH] chunks
        const int r = cidx / NCH, j = cidx % NCH;
        if (row0 + r < nrows)
            *reinterpret_cast<uint4*>(colp + (row0 + r) * 64 + j * VEC) =
                *reinterpret_cast<const uint4*>(stage + r * 64 + ((j ^ (r & (NCH - 1))) * VEC));
    }
    }
}

// wp[(g*128 + j*64 + o)][k], k = (vr*3+kw)*Cin + c < 12*Cin (zero padded to 64), from the OIHW fp32 filter
template <typename T>
__global__ void pack_first_pairs_kernel(const float* __restrict__ w, T* __restrict__ wp, int O, int Cin) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2L * O * 64) return;
    const int k = (int)(i % 64), row = (int)(i / 64);
    const int g = row / 128, j = (row / 64) & 1, o = g * 64 + (row & 63);
    float val = 0.f;
    if (k < 12 * Cin) {
        const int c = k % Cin, v = k / Cin, vr = v / 3, kw = v % 3, kh = vr - j;
        if (kh >= 0 && kh <= 2) val = w[((long)(o * Cin + c) * 3 + kh) * 3 + kw];
    }
    wp[i] = from_float<T>(val);
}

// dW[o][c][kh][kw] = sum_z sum_j ws[z][((kh+j)*3 + kw)*Cin + c][j*O + o]      (ws rows = k (64), columns = (j, o))
__global__ void first_pairs_fold_kernel(const float* __restrict__ ws, int nsplit, int O, int Cin, float* __restrict__ dw) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= O * Cin * 9) return;
    const int kw = i % 3, kh = (i / 3) % 3, c = (i / 9) % Cin, o = i / (9 * Cin);
    float t = 0.f;
    for (int z = 0; z < nsplit; ++z) {
        const float* wz = ws + (long)z * 64 * 2 * O;
#pragma unroll
        for (int j = 0; j < 2; ++j) t += wz[(long)(((kh + j) * 3 + kw) * Cin + c) * 2 * O + j * O + o];
    }
    dw[i] = t;
}

template <typename T>
__global__ void im2col3x3_nchw_kernel(const float* __restrict__ x, T* __restrict__ col, int B, int Cin, int H, int W,
                                      int Kpad) {
    constexpr int VEC = VecTraits<T>::N;
    const int cpr = Kpad / VEC;                                   // 16-byte chunks per im2col row
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)B * H * W * cpr;
    if (i >= total) return;
    const int j = (int)(i % cpr);
    const long p = i / cpr;
    float v[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) v[e] = 0.f;
    if (j * VEC < 9 * Cin) {
        const int w = (int)(p % W), h = (int)((p / W) % H);
        const long b = p / ((long)W * H);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int k = j * VEC + e;
            if (k < 9 * Cin) {
                const int tap = k / Cin, c = k % Cin;
                const int hh = h + tap / 3 - 1, ww = w + tap % 3 - 1;
                if (hh >= 0 && hh < H && ww >= 0 && ww < W) v[e] = __ldg(x + ((b * Cin + c) * H + hh) * W + ww);
            }
        }
    }
    store_vec(col + p * Kpad + j * VEC, v);
}
// NCHW fp32 -> NHWC T (generic; used for inputs with Cin a multiple of the vector width and in tests)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int ld, int B, int C, int H, int W) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)B * C * H * W;
    if (i >= total) return;
    const int c = (int)(i % C);
    const long p = i / C;
    const long hw = p % ((long)H * W), b = p / ((long)H * W);
    y[p * ld + c] = from_float<T>(x[(b * C + c) * H * W + hw]);
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int ld, float* __restrict__ y, int B, int C, int H, int W) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)B * C * H * W;
    if (i >= total) return;
    const long hw = i % ((long)H * W);
    const int c = (int)((i / ((long)H * W)) % C);
    const long b = i / ((long)H * W * C);
    y[i] = to_float(x[(b * H * W + hw) * ld + c]);
}
// Conv2d weight (O,C,3,3) fp32 -> forward operand [O][ldk] with k = tap*C + c (zero padded), and optionally the
// dgrad operand [C][9*O] with k = tap'*O + o, tap' = 8 - tap (180-degree rotated filter).
template <typename T>
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, T* __restrict__ wf, int ldk, T* __restrict__ wd, int O,
                                    int C) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)O * ldk) return;
    const int k = (int)(i % ldk), o = (int)(i / ldk);
    float v = 0.f;
    if (k < 9 * C) {
        const int tap = k / C, c = k % C;
        v = w[((long)o * C + c) * 9 + tap];
        if (wd) wd[(long)c * 9 * O + (long)(8 - tap) * O + o] = from_float<T>(v);
    }
    wf[i] = from_float<T>(v);
}
// The same packing for C, O multiples of 32 through a shared-memory tile (32 o x 32 c x 9 taps): the OIHW reads, the
// [O][tap*C+c] rows and the [C][tap'*O+o] rows are all written as 64..1152-byte runs instead of 2-byte scatters.
template <typename T>
__global__ void __launch_bounds__(256) pack_conv3x3_tiled_kernel(const float* __restrict__ w, T* __restrict__ wf, int ldk,
                                                                 T* __restrict__ wd, int O, int C) {
    __shared__ float sm[32][289];                     // [o][c*9 + tap], row padded against bank conflicts
    const int o0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int idx = threadIdx.x; idx < 32 * 288; idx += 256) {
        const int ol = idx / 288, r = idx % 288;
        sm[ol][r] = w[((long)(o0 + ol) * C + c0) * 9 + r];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * 288; idx += 256) {
        const int cl = idx & 31, tap = (idx >> 5) % 9, ol = idx / 288;
        wf[(long)(o0 + ol) * ldk + tap * C + c0 + cl] = from_float<T>(sm[ol][cl * 9 + tap]);
    }
    if (wd) {
        for (int idx = threadIdx.x; idx < 32 * 288; idx += 256) {
            const int ol = idx & 31, tap = (idx >> 5) % 9, cl = idx / 288;
            wd[(long)(c0 + cl) * 9 * O + (long)(8 - tap) * O + o0 + ol] = from_float<T>(sm[ol][cl * 9 + tap]);
        }
    }
}
// ConvTranspose2d weight (Cin,Cout,2,2) fp32 -> forward operand [4*Cout][Cin] (n = (d*2+e)*Cout + o, k = cin) and
// dgrad operand [Cin][4*Cout] (k = (d*2+e)*Cout + o).
template <typename T>
__global__ void pack_convT_kernel(const float* __restrict__ w, T* __restrict__ wf, T* __restrict__ wd, int Cin, int Cout) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)Cin * Cout * 4) return;
    const int de = (int)(i % 4);
    const int o = (int)((i / 4) % Cout);
    const int ci = (int)(i / (4L * Cout));
    const T v = from_float<T>(w[i]);
    if (wf) wf[((long)de * Cout + o) * Cin + ci] = v;
    if (wd) wd[(long)ci * 4 * Cout + (long)de * Cout + o] = v;
}
// split-K partial reduction with layout change, fp32 out in the parameter's own layout.
//  mode 0: ws[s][O][ldn] with n = tap*C + c  ->  dW (O,C,3,3)
//  mode 1: ws[s][Cin][4*Cout] with n = de*Cout + o -> dW (Cin,Cout,2,2)
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int nsplit, long split_stride, int mode, int D0,
                                    int D1, int ldn, float* __restrict__ dw) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    long src;
    if (mode == 0) {            // D0 = O, D1 = C
        if (i >= (long)D0 * D1 * 9) return;
        const int tap = (int)(i % 9), c = (int)((i / 9) % D1), o = (int)(i / (9L * D1));
        src = (long)o * ldn + (long)tap * D1 + c;
    } else {                    // D0 = Cin, D1 = Cout
        if (i >= (long)D0 * D1 * 4) return;
        const int de = (int)(i % 4), o = (int)((i / 4) % D1), ci = (int)(i / (4L * D1));
        src = (long)ci * ldn + (long)de * D1 + o;
    }
    float t = 0.f;
    for (int s = 0; s < nsplit; ++s) t += ws[(long)s * split_stride + src];
    dw[i] = t;
}

// ---------------------------------------------------------------------------------------------------------
// Standalone SELayer (UCA:61-72 called on its own, NCHW fp32, inputs of any sign): the two plane-wise passes around
// unetca_se_fc / unetca_se_fc_bwd.  One block per (b, c) plane of hw contiguous floats.
//   plane_dot:        out[plane] = sum_i a[i] * (b ? b[i] : 1)          squeeze (b null) and ds = sum dy * x
//   plane_scale_add:  out[i] = x[i] * s[plane] + (t ? t[plane] * tscale : 0)     y = x * s;  dx = dy * s + dp / HW
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) plane_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long hw,
                                                        float* __restrict__ out) {
    const float* pa = a + (long)blockIdx.x * hw;
    const float* pb = b ? b + (long)blockIdx.x * hw : nullptr;
    float acc = 0.f;
    for (long i = threadIdx.x; i < hw; i += 256) acc = pb ? fmaf(pa[i], pb[i], acc) : acc + pa[i];
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        out[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) plane_scale_add_kernel(const float* __restrict__ x, const float* __restrict__ s,
                                                              const float* __restrict__ t, float tscale, long hw,
                                                              float* __restrict__ out) {
    const long base = (long)blockIdx.x * hw;
    const float sv = s[blockIdx.x], tv = t ? t[blockIdx.x] * tscale : 0.f;
    for (long i = threadIdx.x; i < hw; i += 256) out[base + i] = fmaf(x[base + i], sv, tv);
}

}  // namespace unetca

// =========================================================================================================
// C ABI
// =========================================================================================================
using namespace unetca;

#define DISPATCH_T(dtype, ...)                                                          \
    do {                                                                                \
        if ((dtype) == UNETCA_DTYPE_F32) { typedef float T; __VA_ARGS__; }              \
        else if ((dtype) == UNETCA_DTYPE_BF16) { typedef bf16 T; __VA_ARGS__; }         \
        else { set_error("bad dtype %d", (int)(dtype)); return UNETCA_ERR_ARG; }        \
    } while (0)

template <typename T> static bool chan_ok(int C, int ld) {
    const int V = VecTraits<T>::N;
    return C > 0 && C % V == 0 && C / V <= kThreads && ld % V == 0 && ld >= C;
}
#define REQ_CHAN(C, ld) UNETCA_REQUIRE(chan_ok<T>(C, ld), "%s: C=%d ld=%d unsupported (C must be a multiple of %d, <= %d)", \
                                       __func__, C, ld, VecTraits<T>::N, kThreads * VecTraits<T>::N)

// out = relu(a*y+b) [* s[b,c]] over dense bf16 tensors through se_scale_stream_kernel
static int launch_se_scale_stream(const void* y, void* out, int B, long hw, int C, const float* scale, const float* shift,
                                  const float* s, cudaStream_t st) {
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        if (cudaFuncSetAttribute(se_scale_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kScSmemBytes) != cudaSuccess) {
            set_error("se_scale: cudaFuncSetAttribute failed"); return UNETCA_ERR_CUDA;
        }
        attr_done = true;
    }
    long chunk = (long)g_apply_stream * kScTile / (C * 2);
    if (chunk < 1) chunk = 1;
    dim3 grid(ceil_div(hw, chunk), B);
    se_scale_stream_kernel<<<grid, kThreads, kScSmemBytes, st>>>((const bf16*)y, (bf16*)out, C, hw, chunk, scale, shift, s);
    return check_launch("se_scale (stream)");
}

extern "C" {

// capacity (in partial rows) every `parts` scratch argument must provide for a batch of B images
int unetca_max_parts(int B) { return kMaxParts + B; }

int unetca_chan_stats(int dtype, const void* y, int ld, int C, long npix, float* parts, int* nparts, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ld);
        const long chunk = red_chunk<T>(C, npix);
        const int nblk = ceil_div(npix, chunk);
        chan_stats_kernel<T><<<nblk, kThreads, 0, (cudaStream_t)stream>>>((const T*)y, ld, C, npix, chunk, parts);
        *nparts = nblk;
    });
    return check_launch("chan_stats");
}

int unetca_bn_finalize_train(const float* parts, int nparts, int C, long count, const float* conv_bias,
                             const float* gamma, const float* beta, float* running_mean, float* running_var,
                             float momentum, float eps, float* mean, float* invstd, float* scale, float* shift,
                             void* stream) {
    UNETCA_REQUIRE(count > 1, "Expected more than 1 value per channel when training, got %ld", count);
    bn_finalize_train_kernel<<<ceil_div(C, 32), 256, 0, (cudaStream_t)stream>>>(
        parts, nparts, C, (double)count, conv_bias, gamma, beta, running_mean, running_var, momentum, eps, mean, invstd,
        scale, shift);
    return check_launch("bn_finalize_train");
}

int unetca_bn_fold_eval(int C, const float* conv_bias, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale, float* shift, void* stream) {
    bn_fold_eval_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(C, conv_bias, gamma, beta, running_mean,
                                                                           running_var, eps, scale, shift);
    return check_launch("bn_fold_eval");
}

// out = relu(scale*y+shift) (out may be null when only the SE squeeze sums are wanted);
// pool_parts != null: per-image partial sums [B][*nparts][C]
int unetca_bn_relu(int dtype, const void* y, int ldy, void* out, int ldo, int B, long pix_per_img, int C,
                   const float* scale, const float* shift, float* pool_parts, int* nparts, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldy);
        if (out) REQ_CHAN(C, ldo);
        if (out && !pool_parts && stream_ok<T>(C, ldy, ldo)) {
            if (nparts) *nparts = 0;
            return launch_se_scale_stream(y, out, B, pix_per_img, C, scale, shift, nullptr, (cudaStream_t)stream);
        }
        static int slots_dev[kMaxDevices] = {};
        int& slots = slots_dev[device_slot()];
        if (!slots) slots = resident_blocks(bn_relu_kernel<T, false, true>);
        const long chunk = pool_parts ? img_red_chunk<T>(C, pix_per_img, B, slots) : ew_chunk<T>(C, pix_per_img);
        dim3 grid(ceil_div(pix_per_img, chunk), B);
        cudaStream_t st = (cudaStream_t)stream;
        if (out && pool_parts)
            bn_relu_kernel<T, true, true><<<grid, kThreads, 0, st>>>((const T*)y, ldy, (T*)out, ldo, C, pix_per_img, chunk, scale, shift, pool_parts);
        else if (out)
            bn_relu_kernel<T, true, false><<<grid, kThreads, 0, st>>>((const T*)y, ldy, (T*)out, ldo, C, pix_per_img, chunk, scale, shift, nullptr);
        else if (pool_parts)
            bn_relu_kernel<T, false, true><<<grid, kThreads, 0, st>>>((const T*)y, ldy, nullptr, 0, C, pix_per_img, chunk, scale, shift, pool_parts);
        else { set_error("bn_relu: nothing to do"); return UNETCA_ERR_ARG; }
        if (nparts) *nparts = grid.x;
    });
    return check_launch("bn_relu");
}

int unetca_plane_dot(const float* a, const float* b, long nplanes, long hw, float* out, void* stream) {
    UNETCA_REQUIRE(a && out && nplanes > 0 && hw > 0, "plane_dot: bad arguments");
    plane_dot_kernel<<<(unsigned)nplanes, 256, 0, (cudaStream_t)stream>>>(a, b, hw, out);
    return check_launch("plane_dot");
}
int unetca_plane_scale_add(const float* x, const float* s, const float* t, float tscale, long nplanes, long hw, float* out,
                           void* stream) {
    UNETCA_REQUIRE(x && s && out && nplanes > 0 && hw > 0, "plane_scale_add: bad arguments");
    plane_scale_add_kernel<<<(unsigned)nplanes, 256, 0, (cudaStream_t)stream>>>(x, s, t, tscale, hw, out);
    return check_launch("plane_scale_add");
}

int unetca_se_fc(const float* pool_parts, int nparts, int B, int C, int Cr, long hw, const float* w1, const float* w2,
                 float* p, float* z, float* s, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rows = nparts;
    if (nparts > kPartsPreReduce) {
        img_parts_sum_kernel<<<dim3(ceil_div(C, 32), B), 256, 0, st>>>(const_cast<float*>(pool_parts), nparts, C);
        rows = 1;
    }
    se_fc_kernel<1><<<B, 256, (C + Cr) * sizeof(float), st>>>(pool_parts, rows, nparts, C, Cr, 1.f / (float)hw, w1, w2, p, z, s,
                                                             nullptr, nullptr, nullptr, nullptr);
    return check_launch("se_fc");
}

// SE squeeze partial sums (count of active pixels, masked sum of y) per image.  parts: [B * *nparts][2][C]
int unetca_se_squeeze(int dtype, const void* y, int ldy, int B, long pix_per_img, int C, const float* scale,
                      const float* shift, float* parts, int* nparts, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldy);
        if (stream_ok<T>(C, ldy, ldy)) {
            static int sslots_dev[kMaxDevices] = {};
            int& sslots = sslots_dev[device_slot()];
            if (!sslots) sslots = resident_blocks_smem(reduce_stream_kernel<1>, kRdSmemBytes);
            if (sslots > 0) {
                const long chunk = img_red_chunk<T>(C, pix_per_img, B, sslots);
                dim3 grid(ceil_div(pix_per_img, chunk), B);
                reduce_stream_kernel<1><<<grid, kThreads, kRdSmemBytes, (cudaStream_t)stream>>>(nullptr, (const bf16*)y, C, pix_per_img, chunk, scale, shift, nullptr, parts);
                *nparts = grid.x;
                return check_launch("se_squeeze (stream)");
            }
        }
        static int slots_dev[kMaxDevices] = {};
        int& slots = slots_dev[device_slot()];
        if (!slots) slots = resident_blocks(se_squeeze_kernel<T>);
        const long chunk = img_red_chunk<T>(C, pix_per_img, B, slots);
        dim3 grid(ceil_div(pix_per_img, chunk), B);
        se_squeeze_kernel<T><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const T*)y, ldy, C, pix_per_img, chunk, scale, shift, parts);
        *nparts = grid.x;
    });
    return check_launch("se_squeeze");
}

// FC chain on the squeeze partials of unetca_se_squeeze; mean (nullable, eval) = BN batch mean; sums34 (nullable):
// [B][2][C] = (sum m, sum m*(y-mean)) kept for the merged backward reduction
int unetca_se_fc3(const float* parts2, int nparts, int B, int C, int Cr, long hw, const float* w1, const float* w2,
                  const float* scale, const float* shift, const float* mean, float* p, float* z, float* s, float* sums34,
                  void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rows = nparts;
    if (nparts > kPartsPreReduce) {
        img_parts_sum_kernel<<<dim3(ceil_div(2 * C, 32), B), 256, 0, st>>>(const_cast<float*>(parts2), nparts, 2 * C);
        rows = 1;
    }
    se_fc_kernel<3><<<B, 256, (C + Cr) * sizeof(float), st>>>(parts2, rows, nparts, C, Cr, 1.f / (float)hw, w1, w2, p, z, s, sums34,
                                                             scale, shift, mean);
    return check_launch("se_fc3");
}

// o = relu(scale*y+shift) * s[b,c] -> out (stride ldo); pooled/pos non-null: fused MaxPool2d(2)
int unetca_se_scale_pool(int dtype, const void* y, int ldy, void* out, int ldo, void* pooled, int ldp, uint8_t* pos,
                         int B, int H, int W, int C, const float* scale, const float* shift, const float* s,
                         void* stream) {
    UNETCA_REQUIRE(!pooled || (H % 2 == 0 && W % 2 == 0), "se_scale_pool: H, W must be even to pool (got %d x %d)", H, W);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldy); REQ_CHAN(C, ldo);
        cudaStream_t st = (cudaStream_t)stream;
        if (pooled && qp_ok<T>(C, ldy, ldo, ldp, 8)) {
            static int sslots_dev[kMaxDevices] = {};
            int& sslots = sslots_dev[device_slot()];
            if (!sslots) sslots = resident_blocks_smem(se_scale_pool_stream_kernel, kSpSmemBytes);
            if (sslots > 0) {
                const int QW = kThreads / (C / 8);
                SpMaps maps;
                int rc;
                if ((rc = make_tmap_nhwc(&maps.y, y, 2, C, W, H, B, ldy, 2 * QW, 2)) < 0) return rc;
                if ((rc = make_tmap_nhwc(&maps.out, out, 2, C, W, H, B, ldo, 2 * QW, 2)) < 0) return rc;
                if ((rc = make_tmap_nhwc(&maps.pooled, pooled, 2, C, W / 2, H / 2, B, ldp, QW, 1)) < 0) return rc;
                if ((rc = make_tmap_nhwc(&maps.pos, pos, 1, C, W / 2, H / 2, B, C, QW, 1)) < 0) return rc;
                const int ntile_img = (H / 2) * ceil_div(W / 2, QW);
                const int tpb = 16;
                dim3 grid(ceil_div(ntile_img, tpb), B);
                se_scale_pool_stream_kernel<<<grid, kThreads, kSpSmemBytes, st>>>(maps, H, W, C, tpb, scale, shift, s);
                return check_launch("se_scale_pool (stream)");
            }
        }
        if (!pooled && stream_ok<T>(C, ldy, ldo)) return launch_se_scale_stream(y, out, B, (long)H * W, C, scale, shift, s, st);
        if (pooled) {
            REQ_CHAN(C, ldp);
            const long nquad = (long)(H / 2) * (W / 2);
            const long chunk = (long)row_map<T>(C).rows * g_pool_quads;
            dim3 grid(ceil_div(nquad, chunk), B);
            se_scale_pool_kernel<T, true><<<grid, kThreads, 0, st>>>((const T*)y, ldy, (T*)out, ldo, (T*)pooled, ldp, pos, H, W, C, chunk, scale, shift, s);
        } else {
            const long hw = (long)H * W;
            const long chunk = ew_chunk<T>(C, hw);
            dim3 grid(ceil_div(hw, chunk), B);
            se_scale_kernel<T><<<grid, kThreads, 0, st>>>((const T*)y, ldy, (T*)out, ldo, C, hw, chunk, scale, shift, s);
        }
    });
    return check_launch("se_scale_pool");
}

int unetca_maxpool2x2(int dtype, const void* x, int ldx, void* pooled, int ldp, uint8_t* pos, long long* idx64, int B,
                      int H, int W, int C, void* stream) {
    UNETCA_REQUIRE(H >= 2 && W >= 2, "maxpool2x2: input %d x %d too small", H, W);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldx); REQ_CHAN(C, ldp);
        const long nthr = (long)B * (H / 2) * (W / 2) * (C / VecTraits<T>::N);
        maxpool_kernel<T><<<ceil_div(nthr, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const T*)x, ldx, (T*)pooled, ldp, pos, idx64, B, H, W, C);
    });
    return check_launch("maxpool2x2");
}

int unetca_pool_bwd_add(int dtype, const void* skip_grad, int lds, const void* dpooled, int ldp, const uint8_t* pos,
                        void* dx, int ldx, int B, int H, int W, int C, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldp); REQ_CHAN(C, ldx);
        if (skip_grad) REQ_CHAN(C, lds);
        const long nthr = (long)B * (H / 2) * (W / 2) * (C / VecTraits<T>::N);
        if (nthr > 0)
            pool_bwd_add_kernel<T><<<ceil_div(nthr, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const T*)skip_grad, lds, (const T*)dpooled, ldp, pos, (T*)dx, ldx, B, H, W, C);
        if ((H & 1) || (W & 1)) {
            const long nb = (long)B * (((H & 1) ? W : 0) + ((W & 1) ? (H - (H & 1)) : 0)) * (C / VecTraits<T>::N);
            pool_bwd_border_kernel<T><<<ceil_div(nb, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const T*)skip_grad, lds, (T*)dx, ldx, B, H, W, C);
        }
    });
    return check_launch("pool_bwd_add");
}

// resize guard of the decoder (UCA:138-157): (B,h,w,C) -> (B,H,W,C), bilinear, align_corners=False
int unetca_resize_bilinear_fwd(int dtype, const void* x, int ldx, int h, int w, void* out, int ldo, int H, int W, int B,
                               int C, void* stream) {
    UNETCA_REQUIRE(h > 0 && w > 0 && H >= h && W >= w, "resize_bilinear: %dx%d -> %dx%d (the guard only up-samples)", h, w, H, W);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldx); REQ_CHAN(C, ldo);
        const long nthr = (long)B * H * W * (C / VecTraits<T>::N);
        resize_bilinear_fwd_kernel<T><<<ceil_div(nthr, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const T*)x, ldx, h, w, (T*)out, ldo, H, W, B, C, (float)h / (float)H, (float)w / (float)W);
    });
    return check_launch("resize_bilinear_fwd");
}

// adjoint: dout (B,H,W,C) -> dx (B,h,w,C)
int unetca_resize_bilinear_bwd(int dtype, const void* dout, int ldd, int H, int W, void* dx, int ldx, int h, int w, int B,
                               int C, void* stream) {
    UNETCA_REQUIRE(h > 0 && w > 0 && H >= h && W >= w, "resize_bilinear: %dx%d -> %dx%d (the guard only up-samples)", h, w, H, W);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldd); REQ_CHAN(C, ldx);
        const long nthr = (long)B * h * w * (C / VecTraits<T>::N);
        resize_bilinear_bwd_kernel<T><<<ceil_div(nthr, kThreads), kThreads, 0, (cudaStream_t)stream>>>((const T*)dout, ldd, H, W, (T*)dx, ldx, h, w, B, C, (float)h / (float)H, (float)w / (float)W);
    });
    return check_launch("resize_bilinear_bwd");
}

int unetca_se_bwd_reduce(int dtype, const void* dout, int ldd, const void* y, int ldy, int B, long pix_per_img, int C,
                         const float* scale, const float* shift, float* parts, int* nparts, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldd); REQ_CHAN(C, ldy);
        static int slots_dev[kMaxDevices] = {};
        int& slots = slots_dev[device_slot()];
        if (!slots) slots = resident_blocks(se_bwd_reduce_kernel<T>);
        const long chunk = img_red_chunk<T>(C, pix_per_img, B, slots);
        dim3 grid(ceil_div(pix_per_img, chunk), B);
        se_bwd_reduce_kernel<T><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const T*)dout, ldd, (const T*)y, ldy, C, pix_per_img, chunk, scale, shift, parts);
        *nparts = grid.x;
    });
    return check_launch("se_bwd_reduce");
}

int unetca_se_fc_bwd(const float* parts, int nparts, int B, int C, int Cr, const float* w1, const float* w2,
                     const float* p, const float* z, const float* s, float* dpre2, float* dz, float* dp, float* dw1,
                     float* dw2, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    se_fc_bwd_kernel<<<B, 256, (C + Cr) * sizeof(float), st>>>(parts, nparts, C, Cr, w1, w2, z, s, dpre2, dz, dp);
    se_fc_wgrad_kernel<<<ceil_div((long)C * Cr, 256), 256, 0, st>>>(B, C, Cr, dpre2, dz, p, z, dw1, dw2);
    return check_launch("se_fc_bwd");
}

// se_bn_bwd_reduce with the block-output gradient rebuilt per quad from the skip gradient and the pooled gradient
// (H, W even).  parts: [B * *nparts][2][C]
int unetca_se_bn_bwd_reduce_pool(int dtype, const void* sg, int lds, const void* dpooled, int ldp, const uint8_t* pos,
                                 const void* y, int ldy, int B, int H, int W, int C, const float* scale, const float* shift,
                                 const float* mean, float* parts, int* nparts, void* stream) {
    UNETCA_REQUIRE(H % 2 == 0 && W % 2 == 0, "se_bn_bwd_reduce_pool: H, W must be even (got %d x %d)", H, W);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, lds); REQ_CHAN(C, ldp); REQ_CHAN(C, ldy);
        if (qp_ok<T>(C, lds, ldp, ldy, 8)) {
            static int sslots_dev[kMaxDevices] = {};
            int& sslots = sslots_dev[device_slot()];
            if (!sslots) sslots = resident_blocks_smem(se_bn_bwd_pool_stream_kernel<false>, kQpSmemBytes);
            QpMaps maps;
            int rc = sslots > 0 ? qp_setup(&maps, sg, lds, dpooled, ldp, pos, y, ldy, nullptr, 0, B, H, W, C) : 0;
            if (rc < 0) return rc;
            if (rc > 0) {
                const int QW = kThreads / (C / 8);
                const int ntile_img = (H / 2) * ceil_div(W / 2, QW);
                int nblk = sslots / B; if (nblk < 1) nblk = 1; if (nblk > ntile_img) nblk = ntile_img;
                if ((long)nblk * B > kMaxParts) nblk = kMaxParts / B > 0 ? kMaxParts / B : 1;
                const int tpb = ceil_div(ntile_img, nblk);
                dim3 grid(ceil_div(ntile_img, tpb), B);
                se_bn_bwd_pool_stream_kernel<false><<<grid, kThreads, kQpSmemBytes, (cudaStream_t)stream>>>(maps, H, W, C, tpb, 0.f, scale, shift, mean, nullptr, nullptr, nullptr, nullptr, parts);
                *nparts = grid.x;
                return check_launch("se_bn_bwd_reduce_pool (stream)");
            }
        }
        static int slots_dev[kMaxDevices] = {};
        int& slots = slots_dev[device_slot()];
        if (!slots) slots = resident_blocks(se_bn_bwd_pool_kernel<T, false>);
        const long nquad = (long)(H / 2) * (W / 2);
        const long chunk = img_red_chunk<T>(C, nquad, B, slots);
        dim3 grid(ceil_div(nquad, chunk), B);
        se_bn_bwd_pool_kernel<T, false><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const T*)sg, lds, (const T*)dpooled, ldp, pos, (const T*)y, ldy, nullptr, 0, H, W, C, chunk, 0.f, scale, shift, mean, nullptr, nullptr, nullptr, nullptr, parts);
        *nparts = grid.x;
    });
    return check_launch("se_bn_bwd_reduce_pool");
}

// bn_bwd_apply (SE form) with the same on-the-fly block-output gradient
int unetca_bn_bwd_apply_pool(int dtype, const void* sg, int lds, const void* dpooled, int ldp, const uint8_t* pos,
                             const void* y, int ldy, void* dy, int lddy, int B, int H, int W, int C, const float* scale,
                             const float* shift, const float* mean, const float* invstd, const float* s, const float* dp,
                             const float* coef, void* stream) {
    UNETCA_REQUIRE(H % 2 == 0 && W % 2 == 0, "bn_bwd_apply_pool: H, W must be even (got %d x %d)", H, W);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, lds); REQ_CHAN(C, ldp); REQ_CHAN(C, ldy); REQ_CHAN(C, lddy);
        if (qp_ok<T>(C, lds, ldp, ldy, lddy)) {
            static int sslots_dev[kMaxDevices] = {};
            int& sslots = sslots_dev[device_slot()];
            if (!sslots) sslots = resident_blocks_smem(se_bn_bwd_pool_stream_kernel<true>, kQpSmemBytesApply);
            QpMaps maps;
            int rc = sslots > 0 ? qp_setup(&maps, sg, lds, dpooled, ldp, pos, y, ldy, dy, lddy, B, H, W, C) : 0;
            if (rc < 0) return rc;
            if (rc > 0) {
                const int QW = kThreads / (C / 8);
                const int ntile_img = (H / 2) * ceil_div(W / 2, QW);
                const int tpb = 16;
                dim3 grid(ceil_div(ntile_img, tpb), B);
                se_bn_bwd_pool_stream_kernel<true><<<grid, kThreads, kQpSmemBytesApply, (cudaStream_t)stream>>>(maps, H, W, C, tpb, 1.f / (float)((long)H * W), scale, shift, mean, invstd, s, dp, coef, nullptr);
                return check_launch("bn_bwd_apply_pool (stream)");
            }
        }
        const long nquad = (long)(H / 2) * (W / 2);
        const long chunk = (long)row_map<T>(C).rows * g_pool_quads * 2;
        dim3 grid(ceil_div(nquad, chunk), B);
        se_bn_bwd_pool_kernel<T, true><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const T*)sg, lds, (const T*)dpooled, ldp, pos, (const T*)y, ldy, (T*)dy, lddy, H, W, C, chunk, 1.f / (float)((long)H * W), scale, shift, mean, invstd, s, dp, coef, nullptr);
    });
    return check_launch("bn_bwd_apply_pool");
}

// merged SE + ReLU + BN backward stage 1 (see se_bn_bwd_reduce_kernel).  parts: [B * *nparts][2][C]
int unetca_se_bn_bwd_reduce(int dtype, const void* dout, int ldd, const void* y, int ldy, int B, long pix_per_img, int C,
                            const float* scale, const float* shift, const float* mean, float* parts, int* nparts,
                            void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldd); REQ_CHAN(C, ldy);
        if (stream_ok<T>(C, ldd, ldy)) {
            static int sslots_dev[kMaxDevices] = {};
            int& sslots = sslots_dev[device_slot()];
            if (!sslots) sslots = resident_blocks_smem(reduce_stream_kernel<0>, kRdSmemBytes);
            if (sslots > 0) {
                const long chunk = img_red_chunk<T>(C, pix_per_img, B, sslots);
                dim3 grid(ceil_div(pix_per_img, chunk), B);
                reduce_stream_kernel<0><<<grid, kThreads, kRdSmemBytes, (cudaStream_t)stream>>>((const bf16*)dout, (const bf16*)y, C, pix_per_img, chunk, scale, shift, mean, parts);
                *nparts = grid.x;
                return check_launch("se_bn_bwd_reduce (stream)");
            }
        }
        static int slots_dev[kMaxDevices] = {};
        int& slots = slots_dev[device_slot()];
        if (!slots) slots = resident_blocks(se_bn_bwd_reduce_kernel<T>);
        const long chunk = img_red_chunk<T>(C, pix_per_img, B, slots);
        dim3 grid(ceil_div(pix_per_img, chunk), B);
        se_bn_bwd_reduce_kernel<T><<<grid, kThreads, 0, (cudaStream_t)stream>>>((const T*)dout, ldd, (const T*)y, ldy, C, pix_per_img, chunk, scale, shift, mean, parts);
        *nparts = grid.x;
    });
    return check_launch("se_bn_bwd_reduce");
}

// FC chain of the SE backward from the merged sums; sums: [B][4][C] scratch kept for unetca_bn_bwd_finalize_se
int unetca_se_fc_bwd_fused(const float* parts, int nparts, int B, int C, int Cr, const float* w1, const float* w2,
                           const float* p, const float* z, const float* s, const float* scale, const float* shift,
                           const float* mean, const float* sums34, float* sums, float* dpre2, float* dz, float* dp,
                           float* dw1, float* dw2, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    int rows = nparts;
    if (nparts > kPartsPreReduce) {
        img_parts_sum_kernel<<<dim3(ceil_div(2 * C, 32), B), 256, 0, st>>>(const_cast<float*>(parts), nparts, 2 * C);
        rows = 1;
    }
    se_fc_bwd_fused_kernel<<<B, 256, (C + Cr + 8 * Cr) * sizeof(float), st>>>(parts, rows, nparts, C, Cr, w1, w2, z, s, scale, shift,
                                                                              mean, sums34, sums, dpre2, dz, dp);
    se_fc_wgrad_kernel<<<ceil_div((long)C * Cr, 256), 256, 0, st>>>(B, C, Cr, dpre2, dz, p, z, dw1, dw2);
    return check_launch("se_fc_bwd_fused");
}

int unetca_bn_bwd_finalize_se(const float* sums, int B, int C, long count, long pix_per_img, const float* gamma,
                              const float* invstd, const float* s, const float* dp, float* dgamma, float* dbeta,
                              float* coef, void* stream) {
    bn_bwd_finalize_se_kernel<<<ceil_div(C, 32), 256, 0, (cudaStream_t)stream>>>(sums, B, C, (double)count,
                                                                                1.0 / (double)pix_per_img, gamma, invstd,
                                                                                s, dp, dgamma, dbeta, coef);
    return check_launch("bn_bwd_finalize_se");
}

// tuning knobs for sweeps: key 0 = pixels per thread-row of an elementwise block, 1 = waves of a reduction grid,
// 2 = pool quads per thread, 3 = 8 KB tiles per block of the streamed bn_bwd_apply (0: register kernel)
void unetca_set_tuning(int key, int value) {
    if (key == 0 && value > 0) g_ew_px = value;
    if (key == 1) g_red_waves = value;
    if (key == 2 && value > 0) g_pool_quads = value;
    if (key == 3) g_apply_stream = value;
}

// stage 1 of ReLU+BN backward (s/dp null: no SE in front).  parts: [B * *nparts][2][C]
int unetca_bn_bwd_reduce(int dtype, const void* dout, int ldd, const void* y, int ldy, int B, long pix_per_img, int C,
                         const float* scale, const float* shift, const float* mean, const float* invstd,
                         const float* s, const float* dp, float* parts, int* nparts, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldd); REQ_CHAN(C, ldy);
        if (!s && stream_ok<T>(C, ldd, ldy)) {
            static int sslots_dev[kMaxDevices] = {};
            int& sslots = sslots_dev[device_slot()];
            if (!sslots) sslots = resident_blocks_smem(reduce_stream_kernel<0>, kRdSmemBytes);
            if (sslots > 0) {
                const long chunk = img_red_chunk<T>(C, pix_per_img, B, sslots);
                dim3 grid(ceil_div(pix_per_img, chunk), B);
                reduce_stream_kernel<0><<<grid, kThreads, kRdSmemBytes, (cudaStream_t)stream>>>((const bf16*)dout, (const bf16*)y, C, pix_per_img, chunk, scale, shift, mean, parts);
                *nparts = grid.x * B;
                return check_launch("bn_bwd_reduce (stream)");
            }
        }
        static int slots_dev[kMaxDevices] = {};
        int& slots = slots_dev[device_slot()];
        if (!slots) slots = resident_blocks(bn_bwd_kernel<T, true, false>);
        const long chunk = img_red_chunk<T>(C, pix_per_img, B, slots);
        dim3 grid(ceil_div(pix_per_img, chunk), B);
        cudaStream_t st = (cudaStream_t)stream;
        const float ihw = 1.f / (float)pix_per_img;
        if (s)
            bn_bwd_kernel<T, true, false><<<grid, kThreads, 0, st>>>((const T*)dout, ldd, (const T*)y, ldy, nullptr, 0, C, pix_per_img, chunk, ihw, scale, shift, mean, invstd, s, dp, nullptr, parts);
        else
            bn_bwd_kernel<T, false, false><<<grid, kThreads, 0, st>>>((const T*)dout, ldd, (const T*)y, ldy, nullptr, 0, C, pix_per_img, chunk, ihw, scale, shift, mean, invstd, nullptr, nullptr, nullptr, parts);
        *nparts = grid.x * B;
    });
    return check_launch("bn_bwd_reduce");
}

int unetca_bn_bwd_finalize(const float* parts, int nparts, int C, long count, const float* gamma, const float* invstd,
                           float* dgamma, float* dbeta, float* coef, void* stream) {
    bn_bwd_finalize_kernel<<<ceil_div(C, 32), 256, 0, (cudaStream_t)stream>>>(parts, nparts, C, (double)count, gamma,
                                                                             invstd, dgamma, dbeta, coef);
    return check_launch("bn_bwd_finalize");
}

int unetca_bn_bwd_apply(int dtype, const void* dout, int ldd, const void* y, int ldy, void* dy, int lddy, int B,
                        long pix_per_img, int C, const float* scale, const float* shift, const float* mean,
                        const float* invstd, const float* s, const float* dp, const float* coef, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldd); REQ_CHAN(C, ldy); REQ_CHAN(C, lddy);
        cudaStream_t st = (cudaStream_t)stream;
        const float ihw = 1.f / (float)pix_per_img;
        if (stream_ok<T>(C, ldd, ldy) && lddy == C) {
            static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
            if (!attr_done) {
                cudaError_t e = cudaFuncSetAttribute(bn_bwd_apply_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmemBytes);
                if (e == cudaSuccess)
                    e = cudaFuncSetAttribute(bn_bwd_apply_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmemBytes);
                if (e != cudaSuccess) { set_error("bn_bwd_apply: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
                attr_done = true;
            }
            // g_apply_stream tiles of 8 KB per block (whole pixels: 8 KB is a multiple of the pixel size for C <= 4096)
            long chunk = (long)g_apply_stream * kStTile / (C * 2);
            if (chunk < 1) chunk = 1;
            dim3 grid(ceil_div(pix_per_img, chunk), B);
            if (s)
                bn_bwd_apply_stream_kernel<true><<<grid, kThreads, kStSmemBytes, st>>>((const bf16*)dout, (const bf16*)y, (bf16*)dy, C, pix_per_img, chunk, ihw, scale, shift, mean, invstd, s, dp, coef);
            else
                bn_bwd_apply_stream_kernel<false><<<grid, kThreads, kStSmemBytes, st>>>((const bf16*)dout, (const bf16*)y, (bf16*)dy, C, pix_per_img, chunk, ihw, scale, shift, mean, invstd, nullptr, nullptr, coef);
            return check_launch("bn_bwd_apply (stream)");
        }
        const long chunk = 2 * ew_chunk<T>(C, pix_per_img);
        dim3 grid(ceil_div(pix_per_img, chunk), B);
        if (s)
            bn_bwd_kernel<T, true, true><<<grid, kThreads, 0, st>>>((const T*)dout, ldd, (const T*)y, ldy, (T*)dy, lddy, C, pix_per_img, chunk, ihw, scale, shift, mean, invstd, s, dp, coef, nullptr);
        else
            bn_bwd_kernel<T, false, true><<<grid, kThreads, 0, st>>>((const T*)dout, ldd, (const T*)y, ldy, (T*)dy, lddy, C, pix_per_img, chunk, ihw, scale, shift, mean, invstd, nullptr, nullptr, coef, nullptr);
    });
    return check_launch("bn_bwd_apply");
}

// out[c] = sum over pixels of x[p][c]  (parts scratch: unetca_max_parts()*C floats)
int unetca_chan_sum(int dtype, const void* x, int ld, int C, long npix, float* parts, float* out, void* stream) {
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ld);
        const long chunk = red_chunk<T>(C, npix);
        const int nblk = ceil_div(npix, chunk);
        cudaStream_t st = (cudaStream_t)stream;
        chan_sum_kernel<T><<<nblk, kThreads, 0, st>>>((const T*)x, ld, C, npix, chunk, parts);
        sum_parts_kernel<<<ceil_div(C, 32), 256, 0, st>>>(parts, nblk, C, nullptr, out);
    });
    return check_launch("chan_sum");
}

// out[i] = sum_r parts[r*row_stride + i]: e.g. the ConvTranspose bias gradient (UCA:114) from the per-CTA channel sums
// that the dgrad convolution producing the concat gradient leaves in its statistics epilogue (parts + Cl, stride 2*O)
int unetca_sum_rows(const float* parts, int nrows, long row_stride, int n, float* out, void* stream) {
    UNETCA_REQUIRE(parts && out && nrows >= 1 && n >= 1 && row_stride >= n, "sum_rows: bad arguments");
    sum_rows_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(parts, nrows, row_stride, n, out);
    return check_launch("sum_rows");
}

// counts[(nc+1)][nc] (int64): row = label (row nc: labels outside [0,nc) other than ignore_index), column = argmax
// class.  parts: scratch of unetca_max_parts(B) * (nc+1)*nc 8-byte words.
int unetca_confusion_counts(const float* logits, const long long* target, int nc, int B, long HW, long long ignore_index,
                            void* parts, long long* counts, void* stream) {
    UNETCA_REQUIRE(nc >= 1 && nc <= kMaxMetricClasses, "confusion_counts: num_classes %d unsupported (1..%d)", nc, kMaxMetricClasses);
    const long npix = (long)B * HW;
    int nblk = ceil_div(npix, kThreads * 8);
    if (nblk > kMaxParts) nblk = kMaxParts;
    if (nblk < 1) nblk = 1;
    cudaStream_t st = (cudaStream_t)stream;
    confusion_kernel<<<nblk, kThreads, 0, st>>>(logits, target, nc, npix, HW, ignore_index, (unsigned long long*)parts);
    confusion_finalize_kernel<<<1, 128, 0, st>>>((const unsigned long long*)parts, nblk, (nc + 1) * nc, counts);
    return check_launch("confusion_counts");
}

// images/masks: n uint8 pixels each; out: n fp32 = ((x/255) - mean) / std; lab: n int64 = (long)(mask/255)
int unetca_prep_u8(const uint8_t* img, const uint8_t* mask, float* out, long long* lab, long n, float mean, float stdv,
                   void* stream) {
    UNETCA_REQUIRE(n >= 0 && stdv != 0.f, "prep_u8: n=%ld std=%f", n, (double)stdv);
    if (n == 0) return 0;
    prep_u8_kernel<<<ceil_div(n, (long)kThreads * 16), kThreads, 0, (cudaStream_t)stream>>>(img, mask, out, lab, n, mean, stdv);
    return check_launch("prep_u8");
}

// pixel-pair im2col of the NCHW fp32 network input (Cin <= 5, H even): colp [B*(H/2)*W][64]
int unetca_im2col_pairs(int dtype, const float* x, void* colp, int B, int Cin, int H, int W, void* stream) {
    UNETCA_REQUIRE(Cin >= 1 && Cin <= 5 && H % 2 == 0, "im2col_pairs: Cin=%d (1..5), H=%d (even)", Cin, H);
    DISPATCH_T(dtype, {
        const long n = (long)B * (H / 2) * W;
        cudaStream_t st = (cudaStream_t)stream;
        switch (Cin) {
            case 1: im2col_pairs_kernel<T, 1><<<ceil_div(n, 256), 256, 0, st>>>(x, (T*)colp, B, H, W); break;
            case 2: im2col_pairs_kernel<T, 2><<<ceil_div(n, 256), 256, 0, st>>>(x, (T*)colp, B, H, W); break;
            case 3: im2col_pairs_kernel<T, 3><<<ceil_div(n, 256), 256, 0, st>>>(x, (T*)colp, B, H, W); break;
            case 4: im2col_pairs_kernel<T, 4><<<ceil_div(n, 256), 256, 0, st>>>(x, (T*)colp, B, H, W); break;
            default: im2col_pairs_kernel<T, 5><<<ceil_div(n, 256), 256, 0, st>>>(x, (T*)colp, B, H, W); break;
        }
    });
    return check_launch("im2col_pairs");
}

// pair-packed first-conv filter wp [2*O][64] from the nn.Conv2d weight (O, Cin, 3, 3)
int unetca_pack_first_pairs(int dtype, const float* w, void* wp, int O, int Cin, void* stream) {
    UNETCA_REQUIRE(O % 64 == 0 && Cin >= 1 && Cin <= 5, "pack_first_pairs: O=%d Cin=%d", O, Cin);
    DISPATCH_T(dtype, {
        pack_first_pairs_kernel<T><<<ceil_div(2L * O * 64, 256), 256, 0, (cudaStream_t)stream>>>(w, (T*)wp, O, Cin);
    });
    return check_launch("pack_first_pairs");
}

int unetca_first_pairs_fold(const float* ws, int nsplit, int O, int Cin, float* dw, void* stream) {
    first_pairs_fold_kernel<<<ceil_div((long)O * Cin * 9, 256), 256, 0, (cudaStream_t)stream>>>(ws, nsplit, O, Cin, dw);
    return check_launch("first_pairs_fold");
}

static int nc_pad(int nc) { return nc <= 2 ? 2 : nc <= 4 ? 4 : 8; }

int unetca_outc_fwd(int dtype, const void* x, int ldx, int C, const float* w, const float* bias, int nc, float* logits,
                    int B, long HW, void* stream) {
    UNETCA_REQUIRE(nc >= 1 && nc <= 8, "outc: num_classes %d unsupported (1..8)", nc);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldx);
        const long npix = (long)B * HW;
        cudaStream_t st = (cudaStream_t)stream;
        if (stream_ok<T>(C, ldx, ldx) && C == 64 && nc <= 2 && HW % kOcPix == 0) {
            static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
            if (!attr_done) {
                if (cudaFuncSetAttribute(outc_fwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOcFwdSmem) != cudaSuccess) {
                    set_error("outc_fwd: cudaFuncSetAttribute failed"); return UNETCA_ERR_CUDA;
                }
                attr_done = true;
            }
            const long chunk = (long)g_apply_stream * 2 * kOcPix;            // 16 tiles per block by default
            dim3 grid(ceil_div(HW, chunk), B);
            outc_fwd_stream_kernel<<<grid, kThreads, kOcFwdSmem, st>>>((const bf16*)x, w, bias, nc, logits, HW, chunk);
            return check_launch("outc_fwd (stream)");
        }
        int nblk = ceil_div(npix, kThreads);
        if (nblk > 8 * num_sms()) nblk = 8 * num_sms();
        const int ncp = nc_pad(nc);
        const size_t sm = (size_t)ncp * C * sizeof(float);
        if (ncp == 2) outc_fwd_kernel<T, 2><<<nblk, kThreads, sm, st>>>((const T*)x, ldx, C, w, bias, nc, logits, npix, HW);
        else if (ncp == 4) outc_fwd_kernel<T, 4><<<nblk, kThreads, sm, st>>>((const T*)x, ldx, C, w, bias, nc, logits, npix, HW);
        else outc_fwd_kernel<T, 8><<<nblk, kThreads, sm, st>>>((const T*)x, ldx, C, w, bias, nc, logits, npix, HW);
    });
    return check_launch("outc_fwd");
}

// parts scratch: unetca_max_parts() * (8*C + 8) floats
// Fused output head (see se_scale_outc_fwd_stream_kernel): bf16, dense C = 64 tensors, nc <= 2, HW % 128 == 0; anything else
// returns UNETCA_ERR_UNSUPPORTED and the caller runs the separate passes.
static bool head_fusable(int dtype, int ldy, int C, int nc, long HW) {
    return dtype == UNETCA_DTYPE_BF16 && g_apply_stream > 0 && C == 64 && ldy == 64 && nc >= 1 && nc <= 2 && HW % kOcPix == 0;
}
int unetca_se_scale_outc_fwd(int dtype, const void* y, int ldy, int B, long HW, int C, const float* scale, const float* shift,
                             const float* s, const float* w, const float* bias, int nc, float* logits, void* stream) {
    if (!head_fusable(dtype, ldy, C, nc, HW)) { set_error("se_scale_outc_fwd: unsupported shape / dtype"); return UNETCA_ERR_UNSUPPORTED; }
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        if (cudaFuncSetAttribute(se_scale_outc_fwd_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kOcFwdSmem) != cudaSuccess) {
            set_error("se_scale_outc_fwd: cudaFuncSetAttribute failed"); return UNETCA_ERR_CUDA;
        }
        attr_done = true;
    }
    const long chunk = (long)g_apply_stream * 2 * kOcPix;
    dim3 grid(ceil_div(HW, chunk), B);
    se_scale_outc_fwd_stream_kernel<<<grid, kThreads, kOcFwdSmem, (cudaStream_t)stream>>>((const bf16*)y, scale, shift, s, w, bias, nc, logits, HW, chunk);
    return check_launch("se_scale_outc_fwd");
}
// parts_bn: rows as unetca_se_bn_bwd_reduce writes them ([B * *nparts][2][C], *nparts rows per image); parts_oc: scratch of
// B * *nparts * (2*C + 2) floats; dw (nc, C) and db (nc) receive the outc gradients (times *gscale).
int unetca_outc_bn_bwd_reduce(int dtype, const float* g, const float* gscale, const float* w, int nc, const void* y, int ldy,
                              int B, long HW, int C, const float* scale, const float* shift, const float* mean, const float* s,
                              float* parts_bn, int* nparts, float* parts_oc, long parts_oc_floats, float* dw, float* db,
                              void* stream) {
    if (!head_fusable(dtype, ldy, C, nc, HW)) { set_error("outc_bn_bwd_reduce: unsupported shape / dtype"); return UNETCA_ERR_UNSUPPORTED; }
    static int sslots_dev[kMaxDevices] = {};
    int& sslots = sslots_dev[device_slot()];
    if (!sslots) sslots = resident_blocks_smem(outc_bn_bwd_stream_kernel<0, false>, kOcBwdSmem);
    if (sslots <= 0) { set_error("outc_bn_bwd_reduce: kernel does not fit"); return UNETCA_ERR_CUDA; }
    long per_img = sslots / B; if (per_img < 1) per_img = 1;
    const long chunk = ceil_div(ceil_div(HW, per_img), (long)kOcPix) * kOcPix;
    dim3 grid(ceil_div(HW, chunk), B);
    UNETCA_REQUIRE((long)grid.x * B <= kMaxParts + B && (long)grid.x * B * (2 * C + 2) <= parts_oc_floats,
                   "outc_bn_bwd_reduce: scratch too small for %d x %d partial rows", (int)grid.x, B);
    cudaStream_t st = (cudaStream_t)stream;
    outc_bn_bwd_stream_kernel<0, false><<<grid, kThreads, kOcBwdSmem, st>>>(g, gscale, w, nc, (const bf16*)y, nullptr, HW, chunk, 0.f, scale,
                                                                          shift, mean, nullptr, s, nullptr, nullptr, parts_bn, parts_oc);
    outc_bwd_finalize_kernel<<<ceil_div((long)(nc * C + nc) * 32, 128), 128, 0, st>>>(parts_oc, grid.x * B, 2, nc, C, gscale, dw, db);
    *nparts = grid.x;
    return check_launch("outc_bn_bwd_reduce");
}
int unetca_outc_bn_bwd_apply(int dtype, const float* g, const float* gscale, const float* w, int nc, const void* y, int ldy, void* dy,
                             int lddy, int B, long HW, int C, const float* scale, const float* shift, const float* mean,
                             const float* invstd, const float* s, const float* dp, const float* coef, void* stream) {
    if (!head_fusable(dtype, ldy, C, nc, HW) || lddy != C) { set_error("outc_bn_bwd_apply: unsupported shape / dtype"); return UNETCA_ERR_UNSUPPORTED; }
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(outc_bn_bwd_stream_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOcBwdSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(outc_bn_bwd_stream_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kOcBwdSmem);
        if (e != cudaSuccess) { set_error("outc_bn_bwd_apply: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    const long chunk = (long)g_apply_stream * 2 * kOcPix;            // 16 tiles of 128 pixels per block by default
    dim3 grid(ceil_div(HW, chunk), B);
    cudaStream_t st = (cudaStream_t)stream;
    const float ihw = 1.f / (float)HW;
    if (s && dp)
        outc_bn_bwd_stream_kernel<1, true><<<grid, kThreads, kOcBwdSmem, st>>>(g, gscale, w, nc, (const bf16*)y, (bf16*)dy, HW, chunk, ihw, scale,
                                                                             shift, mean, invstd, s, dp, coef, nullptr, nullptr);
    else
        outc_bn_bwd_stream_kernel<1, false><<<grid, kThreads, kOcBwdSmem, st>>>(g, gscale, w, nc, (const bf16*)y, (bf16*)dy, HW, chunk, ihw, scale,
                                                                              shift, mean, invstd, nullptr, nullptr, coef, nullptr, nullptr);
    return check_launch("outc_bn_bwd_apply");
}

int unetca_outc_bwd(int dtype, const float* g, const float* gscale, const void* x, int ldx, void* dx, int lddx, int C,
                    const float* w, int nc, int B, long HW, float* parts, float* dw, float* db, void* stream) {
    UNETCA_REQUIRE(nc >= 1 && nc <= 8, "outc: num_classes %d unsupported (1..8)", nc);
    DISPATCH_T(dtype, {
        REQ_CHAN(C, ldx); REQ_CHAN(C, lddx);
        const long npix = (long)B * HW;
        cudaStream_t st = (cudaStream_t)stream;
        if (stream_ok<T>(C, ldx, lddx) && C == 64 && nc <= 2 && HW % kOcPix == 0) {
            static int sslots_dev[kMaxDevices] = {};
            int& sslots = sslots_dev[device_slot()];
            if (!sslots) sslots = resident_blocks_smem(outc_bwd_stream_kernel, kOcBwdSmem);
            if (sslots > 0) {
                // one wave of resident blocks, whole 128-pixel tiles per block, blocks never straddle an image
                long per_img = sslots / B; if (per_img < 1) per_img = 1;
                long chunk = ceil_div(ceil_div(HW, per_img), (long)kOcPix) * kOcPix;
                dim3 grid(ceil_div(HW, chunk), B);
                if ((long)grid.x * B <= kMaxParts) {
                    outc_bwd_stream_kernel<<<grid, kThreads, kOcBwdSmem, st>>>(g, gscale, (const bf16*)x, (bf16*)dx, w, nc, HW, chunk, parts);
                    outc_bwd_finalize_kernel<<<ceil_div((long)(nc * C + nc) * 32, 128), 128, 0, st>>>(parts, grid.x * B, 2, nc, C, gscale, dw, db);
                    return check_launch("outc_bwd (stream)");
                }
            }
        }
        const long chunk = red_chunk<T>(C, npix);
        const int nblk = ceil_div(npix, chunk);
        const int ncp = nc_pad(nc);
        if (ncp == 2) outc_bwd_kernel<T, 2><<<nblk, kThreads, 0, st>>>(g, gscale, (const T*)x, ldx, (T*)dx, lddx, C, w, nc, npix, HW, chunk, parts);
        else if (ncp == 4) outc_bwd_kernel<T, 4><<<nblk, kThreads, 0, st>>>(g, gscale, (const T*)x, ldx, (T*)dx, lddx, C, w, nc, npix, HW, chunk, parts);
        else outc_bwd_kernel<T, 8><<<nblk, kThreads, 0, st>>>(g, gscale, (const T*)x, ldx, (T*)dx, lddx, C, w, nc, npix, HW, chunk, parts);
        outc_bwd_finalize_kernel<<<ceil_div((long)(nc * C + nc) * 32, 128), 128, 0, st>>>(parts, nblk, ncp, nc, C, gscale, dw, db);
    });
    return check_launch("outc_bwd");
}

// loss_out[0] = mean CE over valid pixels, loss_out[1] = number of valid pixels; gscale_out[0] = upstream / N.
// g (un-normalised dlogits), mask (argmax), target may each be null.  parts scratch: 2*unetca_max_parts() floats.
int unetca_cross_entropy(const float* logits, const long long* target, int nc, int B, long HW, long long ignore_index,
                         const float* upstream, float* g, long long* mask, float* parts, float* loss_out,
                         float* gscale_out, void* stream) {
    const long npix = (long)B * HW;
    int nblk = ceil_div(npix, kThreads * 4);
    if (nblk > kMaxParts) nblk = kMaxParts;
    if (nblk < 1) nblk = 1;
    cudaStream_t st = (cudaStream_t)stream;
    const bool al16 = (((uintptr_t)logits | (uintptr_t)target | (uintptr_t)g | (uintptr_t)mask) & 15) == 0;
    if (nc == 2 && HW % 4 == 0 && al16) {
        // the reference's case (NUM_CLASSES = 2, UCA:24): four pixels per thread, 16-byte accesses — 24 B/px at HBM speed
        nblk = ceil_div(npix / 4, kThreads * 2);
        if (nblk > kMaxParts) nblk = kMaxParts;
        ce2_vec_kernel<<<nblk, kThreads, 0, st>>>(logits, target, npix / 4, HW, ignore_index, g, mask, parts);
    } else
    ce_kernel<<<nblk, kThreads, 0, st>>>(logits, target, nc, npix, HW, ignore_index, g, mask, parts);
    if (target && loss_out) ce_finalize_kernel<<<1, 32, 0, st>>>(parts, nblk, upstream, loss_out, gscale_out);
    return check_launch("cross_entropy");
}

int unetca_im2col3x3_nchw(int dtype, const float* x, void* col, int B, int Cin, int H, int W, int Kpad, void* stream) {
    UNETCA_REQUIRE(Kpad >= 9 * Cin, "im2col: Kpad %d < 9*Cin", Kpad);
    DISPATCH_T(dtype, {
        UNETCA_REQUIRE(Kpad % VecTraits<T>::N == 0, "im2col: Kpad %d must be a multiple of %d", Kpad, VecTraits<T>::N);
        const long npix = (long)B * H * W;
        const long total = npix * (Kpad / VecTraits<T>::N);
        cudaStream_t st = (cudaStream_t)stream;
        if (Cin == 1) im2col3x3_small_kernel<T, 1><<<ceil_div(npix, 256), 256, 0, st>>>(x, (T*)col, B, H, W, Kpad);
        else if (Cin == 3) im2col3x3_small_kernel<T, 3><<<ceil_div(npix, 256), 256, 0, st>>>(x, (T*)col, B, H, W, Kpad);
        else im2col3x3_nchw_kernel<T><<<ceil_div(total, 256), 256, 0, st>>>(x, (T*)col, B, Cin, H, W, Kpad);
    });
    return check_launch("im2col3x3_nchw");
}
int unetca_nchw_to_nhwc(int dtype, const float* x, void* y, int ld, int B, int C, int H, int W, void* stream) {
    DISPATCH_T(dtype, {
        const long total = (long)B * C * H * W;
        nchw_to_nhwc_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x, (T*)y, ld, B, C, H, W);
    });
    return check_launch("nchw_to_nhwc");
}
int unetca_nhwc_to_nchw(int dtype, const void* x, int ld, float* y, int B, int C, int H, int W, void* stream) {
    DISPATCH_T(dtype, {
        const long total = (long)B * C * H * W;
        nhwc_to_nchw_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, ld, y, B, C, H, W);
    });
    return check_launch("nhwc_to_nchw");
}
int unetca_pack_conv3x3_weight(int dtype, const float* w, void* wf, int ldk, void* wd, int O, int C, void* stream) {
    UNETCA_REQUIRE(ldk >= 9 * C, "pack_conv3x3: ldk %d < 9*C", ldk);
    DISPATCH_T(dtype, {
        const long total = (long)O * ldk;
        if (O % 32 == 0 && C % 32 == 0 && ldk == 9 * C)
            pack_conv3x3_tiled_kernel<T><<<dim3(O / 32, C / 32), 256, 0, (cudaStream_t)stream>>>(w, (T*)wf, ldk, (T*)wd, O, C);
        else
            pack_conv3x3_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (T*)wf, ldk, (T*)wd, O, C);
    });
    return check_launch("pack_conv3x3_weight");
}
int unetca_pack_convT_weight(int dtype, const float* w, void* wf, void* wd, int Cin, int Cout, void* stream) {
    DISPATCH_T(dtype, {
        const long total = (long)Cin * Cout * 4;
        pack_convT_kernel<T><<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(w, (T*)wf, (T*)wd, Cin, Cout);
    });
    return check_launch("pack_convT_weight");
}
int unetca_wgrad_reduce(const float* ws, int nsplit, long split_stride, int mode, int D0, int D1, int ldn, float* dw,
                        void* stream) {
    const long total = (long)D0 * D1 * (mode == 0 ? 9 : 4);
    wgrad_reduce_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(ws, nsplit, split_stride, mode, D0, D1,
                                                                               ldn, dw);
    return check_launch("wgrad_reduce");
}

}  // extern "C"
