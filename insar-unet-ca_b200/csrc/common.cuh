// common.cuh — shared helpers for the U-Net-CA B200 kernels (sm_100a only).
//
// Conventions used by every kernel in this directory:
//   * activations are NHWC ("pixel rows"): element (b,h,w,c) lives at ((b*H+h)*W+w)*ld + c, where
//     `ld` (elements) may exceed C so that a tensor can be a channel-slice view of a concat buffer;
//   * T is the activation/compute storage type: float (fp32 parity mode) or __nv_bfloat16;
//   * all accumulation is fp32 (double in the tiny finalize kernels);
//   * every entry point is stream-ordered, allocates nothing, and never synchronises the host.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define UNETCA_OK 0
#define UNETCA_ERR_ARG (-1)
#define UNETCA_ERR_CUDA (-2)
#define UNETCA_ERR_UNSUPPORTED (-3)
#define UNETCA_ERR_WORKSPACE (-4)

#define UNETCA_DTYPE_F32 0
#define UNETCA_DTYPE_BF16 1

namespace unetca {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define UNETCA_REQUIRE(cond, ...)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            ::unetca::set_error(__VA_ARGS__);                       \
            return UNETCA_ERR_ARG;                                  \
        }                                                           \
    } while (0)

typedef __nv_bfloat16 bf16;

template <typename T> struct VecTraits;
template <> struct VecTraits<float> { static constexpr int N = 4; };
template <> struct VecTraits<bf16>  { static constexpr int N = 8; };

// 16-byte vector load -> fp32 registers
__device__ __forceinline__ void load_vec(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_vec(const bf16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i]     = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void store_vec(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store_vec(bf16* p, const float (&v)[8]) {
    uint4 t;
    t.x = pack_bf16x2(v[0], v[1]);
    t.y = pack_bf16x2(v[2], v[3]);
    t.z = pack_bf16x2(v[4], v[5]);
    t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
}
// value as it will read back after being stored as T
__device__ __forceinline__ float round_to(float v, const float*) { return v; }
__device__ __forceinline__ float round_to(float v, const bf16*) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

constexpr int kMaxDevices = 64;
int device_slot();          // current device ordinal (index into per-device caches)
int num_sms();

}  // namespace unetca
