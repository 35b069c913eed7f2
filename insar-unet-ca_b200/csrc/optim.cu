// optim.cu — the optimizer step either side of backward: optim.Adam (UCA:466) / optimizer.step()
// (UCA:346) on the fp32 master weights as multi-tensor kernels, the 3x3-convolution weights emitting their bf16 (or
// fp32) operand copies — forward [O][tap*C+c] and dgrad [C][tap'*O+o] — in the same pass, so that the next forward
// finds the packed filters ready (no repack pass, ~6 launches instead of one per tensor).
//
// Arithmetic (per element, fp32, as torch.optim.Adam without amsgrad / maximize):
//     g  = grad + weight_decay * p
//     m  = m + (1 - beta1) * (g - m)
//     v  = beta2 * v + (1 - beta2) * g * g
//     p  = p - (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
// The step counter and the two bias corrections live in a small device buffer advanced by adam_tick_kernel, so a
// captured CUDA graph of the whole train step replays correctly (no host-side step count baked into the graph).
//
// HBM-bound: 4 fp32 reads + 3 fp32 writes per parameter (+ 2 packed writes for conv filters).
#include "common.cuh"

#include <string.h>

namespace unetca {
namespace {

constexpr int kAdamThreads = 256;
constexpr int kAdamChunk = 4096;          // elements per block of the generic kernel
constexpr int kAdamMaxTensors = 64;       // tensors per launch (table passed by value: 64 * 40 B + 65 * 4 B < 4 KB)
constexpr int kAdamMaxLayers = 24;        // conv layers per launch of the packing kernel

// hyper[0] = t (step count), [1] = lr / (1 - beta1^t), [2] = sqrt(1 - beta2^t), [3] = 1 - beta1, [4] = beta2,
// [5] = eps, [6] = weight_decay, [7] = 1 - beta2.  The hyper-parameters arrive as doubles (Python floats) and the
// corrections are formed in double, as torch.optim.Adam forms them on the host.
__global__ void adam_tick_kernel(float* hyper, double lr, double beta1, double beta2, double eps, double weight_decay) {
    const double t = (double)hyper[0] + 1.0;
    hyper[0] = (float)t;
    hyper[1] = (float)(lr / (1.0 - pow(beta1, t)));
    hyper[2] = (float)sqrt(1.0 - pow(beta2, t));
    hyper[3] = (float)(1.0 - beta1); hyper[4] = (float)beta2; hyper[5] = (float)eps; hyper[6] = (float)weight_decay;
    hyper[7] = (float)(1.0 - beta2);
}

struct AdamH { float step_size, bc2s, omb1, beta2, eps, wd, omb2; };
__device__ __forceinline__ AdamH load_hyper(const float* __restrict__ h) {
    return AdamH{h[1], h[2], h[3], h[4], h[5], h[6], h[7]};
}
__device__ __forceinline__ void adam_update(const AdamH& h, float& p, float g, float& m, float& v) {
    g = fmaf(h.wd, p, g);
    m = fmaf(h.omb1, g - m, m);
    v = fmaf(h.omb2, g * g, h.beta2 * v);
    const float denom = sqrtf(v) / h.bc2s + h.eps;
    p = p - h.step_size * (m / denom);
}

struct AdamTensor { float* p; const float* g; float* m; float* v; long n; };
struct AdamTable {
    AdamTensor t[kAdamMaxTensors];
    int block_start[kAdamMaxTensors + 1];
    int count;
};

__global__ void __launch_bounds__(kAdamThreads) adam_multi_kernel(const __grid_constant__ AdamTable tab,
                                                                  const float* __restrict__ hyper) {
    // which tensor does this block work on?  (binary search over the block prefix table)
    int lo = 0, hi = tab.count - 1;
    const int blk = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab.block_start[mid] <= blk) lo = mid; else hi = mid - 1;
    }
    const AdamTensor& T = tab.t[lo];
    const AdamH h = load_hyper(hyper);
    const long e0 = (long)(blk - tab.block_start[lo]) * kAdamChunk;
    long e1 = e0 + kAdamChunk; if (e1 > T.n) e1 = T.n;
    const bool al = ((((uintptr_t)T.p | (uintptr_t)T.g | (uintptr_t)T.m | (uintptr_t)T.v) & 15) == 0);
    if (al && e1 - e0 == kAdamChunk) {
        for (long e = e0 + threadIdx.x * 4; e < e1; e += kAdamThreads * 4) {
            float4 p = *reinterpret_cast<float4*>(T.p + e), m = *reinterpret_cast<float4*>(T.m + e),
                   v = *reinterpret_cast<float4*>(T.v + e);
            const float4 g = *reinterpret_cast<const float4*>(T.g + e);
            adam_update(h, p.x, g.x, m.x, v.x); adam_update(h, p.y, g.y, m.y, v.y);
            adam_update(h, p.z, g.z, m.z, v.z); adam_update(h, p.w, g.w, m.w, v.w);
            *reinterpret_cast<float4*>(T.p + e) = p;
            *reinterpret_cast<float4*>(T.m + e) = m;
            *reinterpret_cast<float4*>(T.v + e) = v;
        }
    } else {
        for (long e = e0 + threadIdx.x; e < e1; e += kAdamThreads) {
            float p = T.p[e], m = T.m[e], v = T.v[e];
            adam_update(h, p, T.g[e], m, v);
            T.p[e] = p; T.m[e] = m; T.v[e] = v;
        }
    }
}

// 3x3 convolution weights (O, C, 3, 3), O and C multiples of 32: one block per 32 o x 32 c x 9 tile.  The OIHW rows of
// p / g / m / v are read and written as 288-float runs; the stepped values pass through shared memory and leave as the
// two operand layouts in 64-byte (or longer) runs.
struct AdamConv { float* p; const float* g; float* m; float* v; void* wf; void* wd; int O, C; };
struct AdamConvTable {
    AdamConv l[kAdamMaxLayers];
    int block_start[kAdamMaxLayers + 1];
    int count;
};

template <typename T>
__global__ void __launch_bounds__(256) adam_pack_conv3x3_kernel(const __grid_constant__ AdamConvTable tab,
                                                                const float* __restrict__ hyper) {
    __shared__ float sm[32][289];                     // [o][c*9 + tap], row padded against bank conflicts
    int lo = 0, hi = tab.count - 1;
    const int blk = blockIdx.x;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (tab.block_start[mid] <= blk) lo = mid; else hi = mid - 1;
    }
    const AdamConv& L = tab.l[lo];
    const AdamH h = load_hyper(hyper);
    const int O = L.O, C = L.C;
    const int tile = blk - tab.block_start[lo];
    const int ctiles = C / 32;
    const int o0 = (tile / ctiles) * 32, c0 = (tile % ctiles) * 32;
    for (int idx = threadIdx.x; idx < 32 * 288; idx += 256) {
        const int ol = idx / 288, r = idx % 288;
        const long e = ((long)(o0 + ol) * C + c0) * 9 + r;
        float p = L.p[e], m = L.m[e], v = L.v[e];
        adam_update(h, p, L.g[e], m, v);
        L.p[e] = p; L.m[e] = m; L.v[e] = v;
        sm[ol][r] = p;
    }
    __syncthreads();
    T* wf = (T*)L.wf;
    T* wd = (T*)L.wd;
    if (wf) {
        for (int idx = threadIdx.x; idx < 32 * 288; idx += 256) {
            const int cl = idx & 31, tap = (idx >> 5) % 9, ol = idx / 288;
            wf[(long)(o0 + ol) * 9 * C + tap * C + c0 + cl] = from_float<T>(sm[ol][cl * 9 + tap]);
        }
    }
    if (wd) {
        for (int idx = threadIdx.x; idx < 32 * 288; idx += 256) {
            const int ol = idx & 31, tap = (idx >> 5) % 9, cl = idx / 288;
            wd[(long)(c0 + cl) * 9 * O + (long)(8 - tap) * O + o0 + ol] = from_float<T>(sm[ol][cl * 9 + tap]);
        }
    }
}

}  // namespace
}  // namespace unetca

using namespace unetca;

extern "C" {

int unetca_adam_tick(float* hyper, double lr, double beta1, double beta2, double eps, double weight_decay, void* stream) {
    UNETCA_REQUIRE(hyper != nullptr, "adam_tick: null hyper buffer");
    UNETCA_REQUIRE(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0. && weight_decay >= 0.,
                   "adam_tick: invalid hyper-parameters (lr=%g betas=(%g, %g) eps=%g weight_decay=%g)", lr, beta1, beta2, eps,
                   weight_decay);
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(hyper, lr, beta1, beta2, eps, weight_decay);
    return check_launch("adam_tick");
}

// table: host array of ntensors rows {p, g, m, v, numel} (int64 each)
int unetca_adam_step(const long long* table, int ntensors, const float* hyper, void* stream) {
    UNETCA_REQUIRE(ntensors >= 0 && (ntensors == 0 || table != nullptr) && hyper != nullptr, "adam_step: bad arguments");
    int launches = 0;
    for (int first = 0; first < ntensors; first += kAdamMaxTensors) {
        AdamTable tab;
        memset(&tab, 0, sizeof(tab));
        const int cnt = ntensors - first < kAdamMaxTensors ? ntensors - first : kAdamMaxTensors;
        long blocks = 0;
        for (int i = 0; i < cnt; ++i) {
            const long long* row = table + (long)(first + i) * 5;
            UNETCA_REQUIRE(row[0] && row[1] && row[2] && row[3] && row[4] > 0, "adam_step: tensor %d has a null pointer or no elements",
                           first + i);
            tab.t[i] = AdamTensor{(float*)row[0], (const float*)row[1], (float*)row[2], (float*)row[3], (long)row[4]};
            tab.block_start[i] = (int)blocks;
            blocks += (row[4] + kAdamChunk - 1) / kAdamChunk;
            UNETCA_REQUIRE(blocks < (1L << 30), "adam_step: too many elements in one launch");
        }
        tab.block_start[cnt] = (int)blocks;
        tab.count = cnt;
        adam_multi_kernel<<<(unsigned)blocks, kAdamThreads, 0, (cudaStream_t)stream>>>(tab, hyper);
        int rc = check_launch("adam_step");
        if (rc < 0) return rc;
        ++launches;
    }
    return launches;
}

// table: host array of nlayers rows {p, g, m, v, wf, wd, O, C} (int64 each); wf / wd may be 0 (copy not wanted)
int unetca_adam_step_conv3x3(int dtype, const long long* table, int nlayers, const float* hyper, void* stream) {
    UNETCA_REQUIRE(nlayers >= 0 && (nlayers == 0 || table != nullptr) && hyper != nullptr, "adam_step_conv3x3: bad arguments");
    UNETCA_REQUIRE(dtype == UNETCA_DTYPE_F32 || dtype == UNETCA_DTYPE_BF16, "adam_step_conv3x3: bad dtype %d", dtype);
    int launches = 0;
    for (int first = 0; first < nlayers; first += kAdamMaxLayers) {
        AdamConvTable tab;
        memset(&tab, 0, sizeof(tab));
        const int cnt = nlayers - first < kAdamMaxLayers ? nlayers - first : kAdamMaxLayers;
        long blocks = 0;
        for (int i = 0; i < cnt; ++i) {
            const long long* row = table + (long)(first + i) * 8;
            const long O = row[6], C = row[7];
            UNETCA_REQUIRE(row[0] && row[1] && row[2] && row[3], "adam_step_conv3x3: layer %d has a null pointer", first + i);
            UNETCA_REQUIRE(O > 0 && C > 0 && O % 32 == 0 && C % 32 == 0, "adam_step_conv3x3: layer %d: O=%ld C=%ld must be multiples of 32",
                           first + i, O, C);
            tab.l[i] = AdamConv{(float*)row[0], (const float*)row[1], (float*)row[2], (float*)row[3], (void*)row[4], (void*)row[5],
                                (int)O, (int)C};
            tab.block_start[i] = (int)blocks;
            blocks += (O / 32) * (C / 32);
        }
        tab.block_start[cnt] = (int)blocks;
        tab.count = cnt;
        if (dtype == UNETCA_DTYPE_F32)
            adam_pack_conv3x3_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(tab, hyper);
        else
            adam_pack_conv3x3_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(tab, hyper);
        int rc = check_launch("adam_step_conv3x3");
        if (rc < 0) return rc;
        ++launches;
    }
    return launches;
}

}  // extern "C"
