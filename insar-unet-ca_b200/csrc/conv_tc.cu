// conv_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM engine for the dense contractions of U-Net-CA (sm_100a).
//
//   conv3x3 fwd + dgrad    UCA:81,84 / autograd     out[p][n]  = sum_{tap,c} X[p+s(tap)][c] * Wk[n][tap*C+c]
//   ConvTranspose k2s2 fwd UCA:112..121             out[(p,de)][o] = sum_c X[p][c] * Wk[de*Cout+o][c] + b[o]
//   ConvTranspose dgrad                             dX[p][c]   = sum_{de,o} dOut[(p,de)][o] * Wd[c][de*Cout+o]
//   first conv (K=9*Cin)   UCA:81 (via im2col rows) out[p][n]  = sum_k col[p][k] * Wk[n][k]
//   conv3x3 / convT / first-conv wgrad              dW[m][n]   = sum_p A[p][m] * B[p][n]       (K = pixels)
//
// Kernels (all persistent, one CTA per SM, warp-specialised: TMA producer warp(s), one MMA-issuing thread, four
// epilogue warps; fp32 accumulators double-buffered in TMEM; every operand tile is a cp.async.bulk.tensor box of
// [pixels][64 channels] bf16 with SWIZZLE_128B; TMA out-of-bounds zero fill is the conv padding):
//   tc_conv3x3_hpix_kernel   conv3x3 fwd/dgrad, O % 128 == 0: 128 channels x 256 pixels per MMA, nine taps from ONE
//                            haloed activation tile (shifted K-major B descriptors)
//   tc_conv3x3_pixn_kernel   conv3x3 fwd/dgrad, O = 64: (2 output rows x 64 channels) x 256 pixel pairs over 4x3
//                            virtual taps (pair-packed filter, 5-D row-lattice tensor map); also the per-tap P=1 form
//   tc_kernel<N,false>       pixels on M, one box per tap: first conv GEMM, ConvTranspose fwd/dgrad, fallback
//   tc_wgrad3x3_wide_kernel / tc_wgrad3x3_kernel / tc_kernel<N,true>   weight gradients (MN-major descriptors)
// Why the layouts look like this (measured, tools/umma_rate_probe.cu): an SS-mode tcgen05.mma M=128,K=16 costs
// >= ~100 cycles whatever N is, so only N = 256 reaches the tensor rate; and TMA fill writes share the shared-memory
// bandwidth with the MMA's operand reads, so re-using one haloed tile for all taps is what lifts the ~70 % ceiling.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <string.h>

namespace unetca {

// Activation operand with TWO sources: channels [0, xsplit) come from tensor map m1, channels [xsplit, C) from m2 (at
// c0 - xsplit).  A decoder block's first conv reads torch.cat([skip, up], 1) (UCA:140): with the two halves kept as two DENSE
// tensors — the encoder block writes the skip, the transposed conv the upsampled one — nobody writes or reads 128-byte rows
// at a 256-byte stride, and the concat still costs nothing.  xsplit == 0: single source (a multiple of 64 otherwise).
__device__ __forceinline__ void tma_load_x2(const CUtensorMap* m1, const CUtensorMap* m2, int xsplit, uint64_t* bar, void* dst,
                                            int c0, int w, int h, int b) {
    if (xsplit > 0 && c0 >= xsplit) tma_load_4d(m2, bar, dst, c0 - xsplit, w, h, b);
    else tma_load_4d(m1, bar, dst, c0, w, h, b);
}

// ---------------------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) TcParams {
    CUtensorMap mapA[4];
    CUtensorMap mapB[4];
    CUtensorMap mapOut[4];
    int num_m_blocks, num_n_blocks;
    int tilesW, tilesH, nimg;      // pixel tiling: M tiles (forward-like) or K tiles (weight gradient)
    int TW, TH, H, W;
    // forward-like
    int ntaps, cchunks;
    signed char dh[12], dw[12], amap[12];
    int out_chunks_per_map;      // 64-channel chunks per output map (ConvTranspose: one map per (d,e) sub-pixel)
    const float* bias;
    int bias_mod;
    float* stat_parts;
    int N;
    // weight gradient
    int a_chunks, a_cchunks, b_chunks_per_map;
    int nsplit, ktiles_total;
    float* ws;
    long long split_stride;
    int ldn, store_transposed, m_valid;
};

constexpr int kTcThreads = 192;
constexpr int kBoxBytesFwd = 128 * 128;   // [128 pixels][64 bf16]
constexpr int kBoxBytesWg = 64 * 128;     // [64 pixels][64 bf16]

template <int BLOCK_N, bool WGRAD> struct TcCfg {
    static constexpr int A_BYTES = WGRAD ? 2 * kBoxBytesWg : kBoxBytesFwd;
    static constexpr int B_BYTES = WGRAD ? (BLOCK_N / 64) * kBoxBytesWg : BLOCK_N * 128;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BYTES = WGRAD ? 0 : (BLOCK_N / 64) * kBoxBytesFwd;
    // per-CTA BatchNorm partial sums (up to 1024 channels) + bias[BLOCK_N] + per-warp partials [4][2][BLOCK_N]
    static constexpr int STAT_BYTES = WGRAD ? 0 : 2 * 1024 * 4 + BLOCK_N * 4 + 8 * BLOCK_N * 4;
    static constexpr int BUDGET = 227 * 1024 - 1024 /*align slack*/ - 256 /*barriers*/ - OUT_BYTES - STAT_BYTES;
    static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
    static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
    static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + OUT_BYTES + STAT_BYTES + 256;
    static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
};

// ---------------------------------------------------------------------------------------------------------
// forward-like epilogue of one accumulator tile (128 pixels x BLOCK_N channels), executed by the 128 epilogue
// threads: TMEM -> (+bias) -> bf16 -> swizzled staging -> TMA store, plus the per-channel BatchNorm partial sums.
// ---------------------------------------------------------------------------------------------------------
template <int BLOCK_N>
__device__ __forceinline__ void fwd_epilogue_tile(uint32_t t_addr, uint8_t* out_stage, float* sm_bias, float* sm_wpart,
                                                  float* sm_stats, const CUtensorMap* mapOut, int chunks_per_map, int w0, int h0, int b,
                                                  int nb, bool valid, const float* bias, int bias_mod, bool do_stats, int N,
                                                  uint64_t* tempty, int r, int lane, int ep_tid) {
    // the staging buffer must have been drained by the previous tile's TMA store
    if (ep_tid == 0) tma_store_wait_read();
    for (int i = ep_tid; i < BLOCK_N; i += 128)
        sm_bias[i] = bias ? __ldg(bias + (nb * BLOCK_N + i) % bias_mod) : 0.f;
    named_bar_sync(1, 128);
#pragma unroll 1
    for (int j = 0; j < BLOCK_N / 64; ++j) {
        uint32_t v[2][32];
        tmem_ld32(t_addr + j * 64, v[0]);
        tmem_ld32(t_addr + j * 64 + 32, v[1]);
        tmem_wait_ld();
        uint8_t* row = out_stage + j * kBoxBytesFwd + r * 128;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 b0 = *reinterpret_cast<const float4*>(sm_bias + j * 64 + half * 32 + c * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(sm_bias + j * 64 + half * 32 + c * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = valid ? __uint_as_float(v[half][c * 8 + i]) + bb[i] : 0.f;
                uint4 pk;
                pk.x = pack_bf16x2(f[0], f[1]); pk.y = pack_bf16x2(f[2], f[3]);
                pk.z = pack_bf16x2(f[4], f[5]); pk.w = pack_bf16x2(f[6], f[7]);
                const int chunk = half * 4 + c;
                *reinterpret_cast<uint4*>(row + ((chunk ^ (r & 7)) << 4)) = pk;
            }
        }
    }
    tcgen05_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty);
    fence_proxy_async_smem();
    named_bar_sync(1, 128);
    if (ep_tid == 0) {
#pragma unroll 1
        for (int j = 0; j < BLOCK_N / 64; ++j) {
            const int q = nb * (BLOCK_N / 64) + j;          // global 64-channel chunk -> (output map, channel offset)
            tma_store_4d(mapOut + q / chunks_per_map, out_stage + j * kBoxBytesFwd, (q % chunks_per_map) * 64, w0, h0, b);
        }
        tma_store_commit();
    }
    if (do_stats) {
        // per-channel sum / sum-of-squares of the staged bf16 tile: thread -> one 16-byte chunk (8 channels)
        // of every RSTEP-th row; lanes sharing a chunk are folded by shuffles, warps by shared memory.
        constexpr int NCH = BLOCK_N / 8;            // chunks per pixel row over all boxes
        constexpr int RSTEP = 128 / NCH;            // threads per chunk
        const int ch = ep_tid % NCH, r0 = ep_tid / NCH;
        const uint8_t* base = out_stage + (ch >> 3) * kBoxBytesFwd;
        float s1[8], s2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
#pragma unroll 4
        for (int rr = r0; rr < 128; rr += RSTEP) {
            const uint4 t4 = *reinterpret_cast<const uint4*>(base + rr * 128 + (((ch & 7) ^ (rr & 7)) << 4));
            const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float lo = __uint_as_float(w4[i] << 16), hi = __uint_as_float(w4[i] & 0xffff0000u);
                s1[2 * i] += lo; s2[2 * i] = fmaf(lo, lo, s2[2 * i]);
                s1[2 * i + 1] += hi; s2[2 * i + 1] = fmaf(hi, hi, s2[2 * i + 1]);
            }
        }
        if (NCH < 32) {
#pragma unroll
            for (int off = NCH; off < 32; off <<= 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
                    s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], off);
                }
            }
        }
        // one partial per (warp, chunk); for NCH == 32 every lane owns a distinct chunk
        const int ew = ep_tid >> 5;
        if (lane < NCH || NCH >= 32) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sm_wpart[(ew * 2 + 0) * BLOCK_N + ch * 8 + i] = s1[i];
                sm_wpart[(ew * 2 + 1) * BLOCK_N + ch * 8 + i] = s2[i];
            }
        }
        named_bar_sync(1, 128);
        for (int i = ep_tid; i < 2 * BLOCK_N; i += 128) {
            const int which = i / BLOCK_N, col = i % BLOCK_N;
            float tsum = 0.f;
#pragma unroll
            for (int w4i = 0; w4i < 4; ++w4i) tsum += sm_wpart[(w4i * 2 + which) * BLOCK_N + col];
            sm_stats[which * N + nb * BLOCK_N + col] += tsum;
        }
    }
}

template <int BLOCK_N, bool WGRAD>
__global__ void __launch_bounds__(kTcThreads, 1) tc_kernel(const __grid_constant__ TcParams p) {
    using Cfg = TcCfg<BLOCK_N, WGRAD>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* stage_base = smem;
    uint8_t* out_stage = smem + STAGES * Cfg::STAGE_BYTES;
    float* sm_stats = reinterpret_cast<float*>(out_stage + Cfg::OUT_BYTES);
    float* sm_bias = sm_stats + 2048;
    float* sm_wpart = sm_bias + BLOCK_N;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + Cfg::OUT_BYTES + Cfg::STAT_BYTES);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
        fence_barrier_init();
        prefetch_tmap(&p.mapA[0]);
        prefetch_tmap(&p.mapB[0]);
        if (!WGRAD) prefetch_tmap(&p.mapOut[0]);
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    if (!WGRAD && p.stat_parts) {
        for (int i = threadIdx.x; i < 2 * p.N; i += kTcThreads) sm_stats[i] = 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const long num_work = WGRAD ? (long)p.num_m_blocks * p.num_n_blocks * p.nsplit
                                : (long)p.num_m_blocks * p.num_n_blocks;
    const int kt_per_split = WGRAD ? (p.ktiles_total + p.nsplit - 1) / p.nsplit : 0;
    const int kblocks_fwd = p.ntaps * p.cchunks;
    const long mn_items = (long)p.num_m_blocks * p.num_n_blocks;   // weight gradient: the split index is the slow one

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                if (!WGRAD) {
                    const int nb = (int)(t % p.num_n_blocks);
                    const int mt = (int)(t / p.num_n_blocks);
                    const int tw = mt % p.tilesW, th = (mt / p.tilesW) % p.tilesH, b = mt / tiles_per_img;
                    const int w0 = tw * p.TW, h0 = th * p.TH;
                    for (int tap = 0; tap < p.ntaps; ++tap) {
                        const CUtensorMap* ma = &p.mapA[p.amap[tap]];
                        for (int cc = 0; cc < p.cchunks; ++cc) {
                            mbar_wait(&empty_bar[s], ph ^ 1);
                            uint8_t* sa = stage_base + s * Cfg::STAGE_BYTES;
                            mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                            tma_load_4d(ma, &full_bar[s], sa, cc * 64, w0 + p.dw[tap], h0 + p.dh[tap], b);
                            tma_load_4d(&p.mapB[0], &full_bar[s], sa + Cfg::A_BYTES, (tap * p.cchunks + cc) * 64,
                                        nb * BLOCK_N, 0, 0);
                            if (++s == STAGES) { s = 0; ph ^= 1; }
                        }
                    }
                } else {
                    const int z = (int)(t / mn_items);
                    const int nb = (int)((t % mn_items) % p.num_n_blocks);
                    const int mb = (int)((t % mn_items) / p.num_n_blocks);
                    int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                    if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                    for (int kt = kt0; kt < kt1; ++kt) {
                        const int tw = kt % p.tilesW, th = (kt / p.tilesW) % p.tilesH, b = kt / tiles_per_img;
                        const int w0 = tw * p.TW, h0 = th * p.TH;
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        uint8_t* sa = stage_base + s * Cfg::STAGE_BYTES;
                        mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            int q = mb * 2 + i; if (q >= p.a_chunks) q = p.a_chunks - 1;
                            const int tap = q / p.a_cchunks, cc = q % p.a_cchunks;
                            tma_load_4d(&p.mapA[p.amap[tap]], &full_bar[s], sa + i * kBoxBytesWg, cc * 64,
                                        w0 + p.dw[tap], h0 + p.dh[tap], b);
                        }
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j) {
                            const int q = nb * (BLOCK_N / 64) + j;
                            tma_load_4d(&p.mapB[q / p.b_chunks_per_map], &full_bar[s],
                                        sa + Cfg::A_BYTES + j * kBoxBytesWg, (q % p.b_chunks_per_map) * 64, w0, h0, b);
                        }
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, BLOCK_N, WGRAD ? 1 : 0, WGRAD ? 1 : 0);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                int nk;
                if (!WGRAD) nk = kblocks_fwd;
                else {
                    const int z = (int)(t / mn_items);
                    int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                    if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                    nk = kt1 - kt0;
                }
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BLOCK_N;
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(stage_base + s * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint64_t da, db;
                        if (!WGRAD) {
                            da = make_smem_desc(sa + k * 32, 16, 1024);
                            db = make_smem_desc(sb + k * 32, 16, 1024);
                        } else {
                            da = make_smem_desc(sa + k * 2048, kBoxBytesWg, 1024);
                            db = make_smem_desc(sb + k * 2048, kBoxBytesWg, 1024);
                        }
                        umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (nk == 0) {
                    // nothing accumulated: the epilogue treats the tile as zero (flag via tfull arrive only)
                }
                umma_commit(&tfull_bar[as]);
                as ^= 1; if (as == 0) aph ^= 1;
            }
        }
    } else {
        // ================================ epilogue (warps 2..5) ================================
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;             // accumulator row owned by this thread
        const int ep_tid = threadIdx.x - 64;     // 0..127
        int as = 0; uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            mbar_wait(&tfull_bar[as], aph);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * BLOCK_N;
            if (!WGRAD) {
                const int nb = (int)(t % p.num_n_blocks);
                const int mt = (int)(t / p.num_n_blocks);
                const int tw = mt % p.tilesW, th = (mt / p.tilesW) % p.tilesH, b = mt / tiles_per_img;
                const int w0 = tw * p.TW, h0 = th * p.TH;
                const bool valid = (h0 + r / p.TW) < p.H && (w0 + r % p.TW) < p.W;
                fwd_epilogue_tile<BLOCK_N>(t_addr, out_stage, sm_bias, sm_wpart, sm_stats, p.mapOut,
                                           p.out_chunks_per_map, w0, h0, b, nb, valid, p.bias, p.bias_mod,
                                           p.stat_parts != nullptr, p.N, &tempty_bar[as], r, lane, ep_tid);
            } else {
                const int z = (int)(t / mn_items);
                const int nb = (int)((t % mn_items) % p.num_n_blocks);
                const int mb = (int)((t % mn_items) / p.num_n_blocks);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                const bool have = kt1 > kt0;
                const int m = mb * 128 + r;
                const bool mvalid = m < p.m_valid;
                float* wsz = p.ws + (long long)z * p.split_stride;
#pragma unroll
                for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + c32 * 32, v);
                    tmem_wait_ld();
                    const int n0 = nb * BLOCK_N + c32 * 32;
                    if (mvalid) {
                        if (p.store_transposed) {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                wsz[(long long)(n0 + i) * p.ldn + m] = have ? __uint_as_float(v[i]) : 0.f;
                        } else {
                            float4* dst = reinterpret_cast<float4*>(wsz + (long long)m * p.ldn + n0);
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                dst[i] = have ? make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[as]);
            }
            as ^= 1; if (as == 0) aph ^= 1;
        }
        if (!WGRAD) {
            if (ep_tid == 0) tma_store_wait_all();
            named_bar_sync(1, 128);
            if (p.stat_parts) {
                float* dst = p.stat_parts + (long)blockIdx.x * 2 * p.N;
                for (int i = ep_tid; i < 2 * p.N; i += 128) dst[i] = sm_stats[i];
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 weight gradient with tap reuse from ONE haloed activation tile                        (UCA:345)
//   dW[o][tap][c] = sum_p dY[p][o] * X[p + s(tap)][c]
// A 16x8-pixel tile of dY ([128 px][64 o], 16 KB) and the 18x10 haloed tile of X ([180 px][64 c], 23 KB) are
// loaded ONCE per k-tile; all 9 taps are then issued from that single X tile by shifting the start address of
// the MN-major descriptor by whole 128-byte pixel rows ((kh*18 + kw) rows) — the tcgen05 SWIZZLE_128B XOR is a
// function of the absolute shared-memory address, so any row shift (and any LBO/SBO that is a multiple of 128 B)
// addresses the data TMA wrote (verified on B200 by tools/umma_shift_probe.cu, profiles/r1_umma_shift_probe.txt).
// The M=128 rows of one MMA are TWO taps x 64 channels: the second 64-row block is the same tile shifted by
// LBO = (s(tap') - s(tap)) * 128 bytes.  Five such M-blocks ((0,1),(2,3),(4,5),(6,7),(7,8)) x N=64 output channels
// = 320 TMEM columns hold the whole 9-tap gradient of a (64 c, 64 o) block while the CTA streams its pixel range.
// Shared-memory fill drops from 9 boxes per tap-set to 1.4, which is what bounded the generic kernel (L2->SM).
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) Wg3Params {
    CUtensorMap mapX, mapDY;
    CUtensorMap mapX2;           // second activation source (channels >= xsplit), see tma_load_x2
    int xsplit;
    int cchunks, oblocks;
    int tilesW, tilesH, nimg;
    int ktiles_total, nsplit;
    float* ws;
    long long split_stride;
    int ldn, C;
};
constexpr int kWgTW = 16, kWgTH = 8, kWgPitch = kWgTW + 2;
constexpr int kWgXBox = kWgPitch * (kWgTH + 2) * 128;      // 23040 bytes actually written by TMA
constexpr int kWgXBytes = 24 * 1024;                       // slot (keeps every tile 1024-byte aligned)
constexpr int kWgYBytes = kWgTW * kWgTH * 128;             // 16 KB
constexpr int kWgStageBytes = kWgXBytes + kWgYBytes;
constexpr int kWgStages = 5;
constexpr int kWgSmemBytes = 1024 + kWgStages * kWgStageBytes + 256;

__global__ void __launch_bounds__(kTcThreads, 1) tc_wgrad3x3_kernel(const __grid_constant__ Wg3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * kWgStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWgStages;
    uint64_t* tfull_bar = bars + 2 * kWgStages;
    uint64_t* tempty_bar = bars + 2 * kWgStages + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgStages + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWgStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapDY);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const long items = (long)p.cchunks * p.oblocks;
    const long num_work = items * p.nsplit;
    const int kt_per_split = (p.ktiles_total + p.nsplit - 1) / p.nsplit;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items), item = (int)(t % items);
                const int ob = item % p.oblocks, cc = item / p.oblocks;
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                for (int kt = kt0; kt < kt1; ++kt) {
                    const int tw = kt % p.tilesW, th = (kt / p.tilesW) % p.tilesH, b = kt / tiles_per_img;
                    const int w0 = tw * kWgTW, h0 = th * kWgTH;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* xs = smem + s * kWgStageBytes;
                    mbar_expect_tx(&full_bar[s], kWgXBox + kWgYBytes);
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &full_bar[s], xs, cc * 64, w0 - 1, h0 - 1, b);
                    tma_load_4d(&p.mapDY, &full_bar[s], xs + kWgXBytes, ob * 64, w0, h0, b);
                    if (++s == kWgStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
            int s = 0; uint32_t ph = 0, aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                mbar_wait(tempty_bar, aph ^ 1);
                tcgen05_fence_after();
                for (int kt = kt0; kt < kt1; ++kt) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xs = smem_u32(smem + s * kWgStageBytes);
                    const uint32_t ys = xs + kWgXBytes;
#pragma unroll 1
                    for (int hh = 0; hh < kWgTH; ++hh) {
                        const uint64_t db = make_smem_desc(ys + hh * (kWgTW * 128), 8192, 1024);
                        const uint32_t xrow = xs + hh * (kWgPitch * 128);
                        const uint32_t acc = (kt > kt0 || hh > 0) ? 1u : 0u;
                        // tap pairs (0,1) (2,3) (4,5) (6,7) (7,8): first-tap row shift kh*18+kw, LBO = shift difference
                        umma_bf16(tmem_base + 0 * 64, make_smem_desc(xrow + (0 * kWgPitch + 0) * 128, 128, 1024), db, idesc, acc);
                        umma_bf16(tmem_base + 1 * 64, make_smem_desc(xrow + (0 * kWgPitch + 2) * 128, (kWgPitch - 2) * 128, 1024), db, idesc, acc);
                        umma_bf16(tmem_base + 2 * 64, make_smem_desc(xrow + (1 * kWgPitch + 1) * 128, 128, 1024), db, idesc, acc);
                        umma_bf16(tmem_base + 3 * 64, make_smem_desc(xrow + (2 * kWgPitch + 0) * 128, 128, 1024), db, idesc, acc);
                        umma_bf16(tmem_base + 4 * 64, make_smem_desc(xrow + (2 * kWgPitch + 1) * 128, 128, 1024), db, idesc, acc);
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kWgStages) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull_bar);
                aph ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int z = (int)(t / items), item = (int)(t % items);
            const int ob = item % p.oblocks, cc = item / p.oblocks;
            int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
            if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
            const bool have = kt1 > kt0;
            mbar_wait(tfull_bar, aph);
            tcgen05_fence_after();
            float* wsz = p.ws + (long long)z * p.split_stride + (long long)(ob * 64) * p.ldn + cc * 64 + (r & 63);
#pragma unroll 1
            for (int mbk = 0; mbk < 5; ++mbk) {
                const int tap = (mbk < 4 ? 2 * mbk : 7) + (r >> 6);
                const bool rvalid = !(mbk == 4 && r < 64);
                float* dst = wsz + tap * p.C;
#pragma unroll
                for (int c32 = 0; c32 < 2; ++c32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + mbk * 64 + c32 * 32, v);
                    tmem_wait_ld();
                    if (rvalid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            dst[(long long)(c32 * 32 + i) * p.ldn] = have ? __uint_as_float(v[i]) : 0.f;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            aph ^= 1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 weight gradient, wide variant (C and O multiples of 128): one work item = (128 c, 128 o, filter row kh).
// M = 128 = two 64-channel X chunks (LBO = chunk tile), N = 128 = two dY chunks, the three taps kw = 0..2 of filter
// row kh are three 128-column accumulators fed from the same 18x8 haloed X tiles by row-shifted descriptors.
// Shared-memory operand reads per MMA: 4 KB (A) + 4 KB (B) per 64 cycles = 128 B/cycle — the SS-mode limit, versus
// 192 B/cycle needed by the N=64 variant above (which therefore tops out at 2/3 of the MMA rate).
// ---------------------------------------------------------------------------------------------------------
constexpr int kWg2XBytes = kWgPitch * kWgTH * 128;          // 18 KB per 64-channel chunk (no vertical halo needed)
constexpr int kWg2StageBytes = 2 * kWg2XBytes + 2 * kWgYBytes;
constexpr int kWg2Stages = 3;
constexpr int kWg2SmemBytes = 1024 + kWg2Stages * kWg2StageBytes + 256;

__global__ void __launch_bounds__(kTcThreads, 1) tc_wgrad3x3_wide_kernel(const __grid_constant__ Wg3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWg2Stages * kWg2StageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWg2Stages;
    uint64_t* tfull_bar = bars + 2 * kWg2Stages;
    uint64_t* tempty_bar = bars + 2 * kWg2Stages + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWg2Stages + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWg2Stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapDY);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const int cpairs = p.cchunks / 2, opairs = p.oblocks / 2;
    const long items = (long)cpairs * opairs * 3;
    const long num_work = items * p.nsplit;
    const int kt_per_split = (p.ktiles_total + p.nsplit - 1) / p.nsplit;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items), item = (int)(t % items);
                const int kh = item % 3, op = (item / 3) % opairs, cp = item / (3 * opairs);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                for (int kt = kt0; kt < kt1; ++kt) {
                    const int tw = kt % p.tilesW, th = (kt / p.tilesW) % p.tilesH, b = kt / tiles_per_img;
                    const int w0 = tw * kWgTW, h0 = th * kWgTH;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* xs = smem + s * kWg2StageBytes;
                    mbar_expect_tx(&full_bar[s], kWg2StageBytes);
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &full_bar[s], xs, cp * 128, w0 - 1, h0 - 1 + kh, b);
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &full_bar[s], xs + kWg2XBytes, cp * 128 + 64, w0 - 1, h0 - 1 + kh, b);
                    tma_load_4d(&p.mapDY, &full_bar[s], xs + 2 * kWg2XBytes, op * 128, w0, h0, b);
                    tma_load_4d(&p.mapDY, &full_bar[s], xs + 2 * kWg2XBytes + kWgYBytes, op * 128 + 64, w0, h0, b);
                    if (++s == kWg2Stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 128, 1, 1);
            int s = 0; uint32_t ph = 0, aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                mbar_wait(tempty_bar, aph ^ 1);
                tcgen05_fence_after();
                for (int kt = kt0; kt < kt1; ++kt) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xs = smem_u32(smem + s * kWg2StageBytes);
                    const uint32_t ys = xs + 2 * kWg2XBytes;
#pragma unroll 1
                    for (int hh = 0; hh < kWgTH; ++hh) {
                        const uint64_t db = make_smem_desc(ys + hh * (kWgTW * 128), kWgYBytes, 1024);
                        const uint32_t xrow = xs + hh * (kWgPitch * 128);
                        const uint32_t acc = (kt > kt0 || hh > 0) ? 1u : 0u;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            umma_bf16(tmem_base + kw * 128, make_smem_desc(xrow + kw * 128, kWg2XBytes, 1024), db, idesc, acc);
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kWg2Stages) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull_bar);
                aph ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int z = (int)(t / items), item = (int)(t % items);
            const int kh = item % 3, op = (item / 3) % opairs, cp = item / (3 * opairs);
            int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
            if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
            const bool have = kt1 > kt0;
            mbar_wait(tfull_bar, aph);
            tcgen05_fence_after();
            float* wsz = p.ws + (long long)z * p.split_stride + (long long)(op * 128) * p.ldn + cp * 128 + r;
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
                float* dst = wsz + (kh * 3 + kw) * p.C;
#pragma unroll 1
                for (int c32 = 0; c32 < 4; ++c32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + kw * 128 + c32 * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        dst[(long long)(c32 * 32 + i) * p.ldn] = have ? __uint_as_float(v[i]) : 0.f;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            aph ^= 1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}


// ---------------------------------------------------------------------------------------------------------
// conv3x3 weight gradient for O % 256 == 0 (and C % 128 == 0): one work item = (128 c, 256 o, tap pair).
// The wide kernel above sits at the MN-major A-fetch floor (91 cycles per 128x128x16 MMA, 70 % of the tensor rate).
// TMEM's 512 columns hold exactly TWO 256-column accumulators, so here the nine taps are walked as the pairs
// (0,1) (2,3) (4,5) (6,7) and the single 8: every MMA is 128x256x16 (full rate).  Both taps of a pair come from one
// 18 x 9 haloed X tile per 64-channel chunk (a pair spans at most two filter rows) by row/column-shifted descriptors.
// ---------------------------------------------------------------------------------------------------------
constexpr int kWg3XRows = kWgTH + 1;
constexpr int kWg3XBox = kWgPitch * kWg3XRows * 128;        // 20736 bytes written by TMA per chunk
constexpr int kWg3XBytes = 21 * 1024;                       // slot
constexpr int kWg3StageBytes = 2 * kWg3XBytes + 4 * kWgYBytes;
constexpr int kWg3Stages = 2;
constexpr int kWg3SmemBytes = 1024 + kWg3Stages * kWg3StageBytes + 256;
static_assert(kWg3SmemBytes <= 227 * 1024, "wide256 wgrad: shared memory budget");

__global__ void __launch_bounds__(kTcThreads, 1) tc_wgrad3x3_wide256_kernel(const __grid_constant__ Wg3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWg3Stages * kWg3StageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWg3Stages;
    uint64_t* tfull_bar = bars + 2 * kWg3Stages;
    uint64_t* tempty_bar = bars + 2 * kWg3Stages + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWg3Stages + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWg3Stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapDY);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const int cpairs = p.cchunks / 2, oquads = p.oblocks / 4;
    const long items = (long)cpairs * oquads * 5;
    const long num_work = items * p.nsplit;
    const int kt_per_split = (p.ktiles_total + p.nsplit - 1) / p.nsplit;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items), item = (int)(t % items);
                const int tp = item % 5, oq = (item / 5) % oquads, cp = item / (5 * oquads);
                const int kh0 = (2 * tp) / 3;
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                for (int kt = kt0; kt < kt1; ++kt) {
                    const int tw = kt % p.tilesW, th = (kt / p.tilesW) % p.tilesH, b = kt / tiles_per_img;
                    const int w0 = tw * kWgTW, h0 = th * kWgTH;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* xs = smem + s * kWg3StageBytes;
                    mbar_expect_tx(&full_bar[s], 2 * kWg3XBox + 4 * kWgYBytes);
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &full_bar[s], xs, cp * 128, w0 - 1, h0 - 1 + kh0, b);
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &full_bar[s], xs + kWg3XBytes, cp * 128 + 64, w0 - 1, h0 - 1 + kh0, b);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        tma_load_4d(&p.mapDY, &full_bar[s], xs + 2 * kWg3XBytes + j * kWgYBytes, oq * 256 + j * 64, w0, h0, b);
                    if (++s == kWg3Stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 1, 1);
            int s = 0; uint32_t ph = 0, aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items), item = (int)(t % items);
                const int tp = item % 5;
                const int t0 = 2 * tp, ntap = tp < 4 ? 2 : 1;
                const int kh0 = t0 / 3;
                // shift of each tap inside the haloed tile (rows relative to kh0)
                const uint32_t off0 = (uint32_t)((t0 % 3) * 128);
                const uint32_t off1 = (uint32_t)((((t0 + 1) / 3 - kh0) * kWgPitch + (t0 + 1) % 3) * 128);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                mbar_wait(tempty_bar, aph ^ 1);
                tcgen05_fence_after();
                for (int kt = kt0; kt < kt1; ++kt) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xs = smem_u32(smem + s * kWg3StageBytes);
                    const uint32_t ys = xs + 2 * kWg3XBytes;
#pragma unroll 1
                    for (int hh = 0; hh < kWgTH; ++hh) {
                        const uint64_t db = make_smem_desc(ys + hh * (kWgTW * 128), kWgYBytes, 1024);
                        const uint32_t xrow = xs + hh * (kWgPitch * 128);
                        const uint32_t acc = (kt > kt0 || hh > 0) ? 1u : 0u;
                        umma_bf16(tmem_base, make_smem_desc(xrow + off0, kWg3XBytes, 1024), db, idesc, acc);
                        if (ntap == 2)
                            umma_bf16(tmem_base + 256, make_smem_desc(xrow + off1, kWg3XBytes, 1024), db, idesc, acc);
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kWg3Stages) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull_bar);
                aph ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int z = (int)(t / items), item = (int)(t % items);
            const int tp = item % 5, oq = (item / 5) % oquads, cp = item / (5 * oquads);
            const int ntap = tp < 4 ? 2 : 1;
            int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
            if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
            const bool have = kt1 > kt0;
            mbar_wait(tfull_bar, aph);
            tcgen05_fence_after();
            float* wsz = p.ws + (long long)z * p.split_stride + (long long)(oq * 256) * p.ldn + cp * 128 + r;
#pragma unroll 1
            for (int j = 0; j < ntap; ++j) {
                float* dst = wsz + (2 * tp + j) * p.C;
#pragma unroll 1
                for (int c32 = 0; c32 < 8; ++c32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + j * 256 + c32 * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        dst[(long long)(c32 * 32 + i) * p.ldn] = have ? __uint_as_float(v[i]) : 0.f;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            aph ^= 1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 weight gradient when C or O is only a multiple of 64 (the 512^2 / 256^2 layers): row-pair layout.
// The kernel above (M = two taps x 64 c, N = 64 o) sits at the MN-major A-fetch floor: 91 cycles for a 32-cycle MMA.
// Here one MMA is  D[(j, o)][(kw, c)]  with M = 128 = 2 dY row-shifts x 64 o and N = 192 = 3 kw taps x 64 c:
//   A block j = the dY tile row hh + j (MN-major, LBO = one 16-pixel tile row), K = the 16 pixels of tile row hh;
//   B block kw = the X view shifted by kw pixels (LBO = 128 B) at X row hh + v, v = 0..3 the "view" (row shift).
// Block j = 0 accumulates tap kh = v, block j = 1 accumulates tap kh = v - 1 (dY row hh+1 against the same X row), so
// walking only the EVEN tile rows hh covers every dY row exactly once: even rows through the j = 0 halves, odd rows
// through the j = 1 halves.  6 of the 8 (view, j) halves are real taps (75 % useful work at N = 192, 96 cycles per MMA,
// instead of 35 %).  The two halves of a tap land in different accumulators, so they are written as two split-K
// slices (2z + j) and summed by the existing reduction.  One work item = (64 c, 64 o, view pair); 384 TMEM columns.
// ---------------------------------------------------------------------------------------------------------
constexpr int kWg4XRows = kWgTH + 1;
constexpr int kWg4XBox = kWgPitch * kWg4XRows * 128;        // 20736
constexpr int kWg4XBytes = 21 * 1024;
constexpr int kWg4StageBytes = kWg4XBytes + kWgYBytes;      // 37 KB
constexpr int kWg4Stages = 5;
constexpr int kWg4SmemBytes = 1024 + kWg4Stages * kWg4StageBytes + 256;
static_assert(kWg4SmemBytes <= 227 * 1024, "row-pair wgrad: shared memory budget");

__global__ void __launch_bounds__(kTcThreads, 1) tc_wgrad3x3_rowpair_kernel(const __grid_constant__ Wg3Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWg4Stages * kWg4StageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kWg4Stages;
    uint64_t* tfull_bar = bars + 2 * kWg4Stages;
    uint64_t* tempty_bar = bars + 2 * kWg4Stages + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWg4Stages + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < kWg4Stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1);
        mbar_init(tempty_bar, 4);
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapDY);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const long items = (long)p.cchunks * p.oblocks * 2;
    const long num_work = items * p.nsplit;
    const int kt_per_split = (p.ktiles_total + p.nsplit - 1) / p.nsplit;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items), item = (int)(t % items);
                const int vp = item & 1, ob = (item >> 1) % p.oblocks, cc = (item >> 1) / p.oblocks;
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                for (int kt = kt0; kt < kt1; ++kt) {
                    const int tw = kt % p.tilesW, th = (kt / p.tilesW) % p.tilesH, b = kt / tiles_per_img;
                    const int w0 = tw * kWgTW, h0 = th * kWgTH;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* xs = smem + s * kWg4StageBytes;
                    mbar_expect_tx(&full_bar[s], kWg4XBox + kWgYBytes);
                    // X rows h0 - 1 + 2*vp ... (+9): both views of the pair
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &full_bar[s], xs, cc * 64, w0 - 1, h0 - 1 + 2 * vp, b);
                    tma_load_4d(&p.mapDY, &full_bar[s], xs + kWg4XBytes, ob * 64, w0, h0, b);
                    if (++s == kWg4Stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 192, 1, 1);
            int s = 0; uint32_t ph = 0, aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / items);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                mbar_wait(tempty_bar, aph ^ 1);
                tcgen05_fence_after();
                for (int kt = kt0; kt < kt1; ++kt) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xs = smem_u32(smem + s * kWg4StageBytes);
                    const uint32_t ys = xs + kWg4XBytes;
#pragma unroll 1
                    for (int hh = 0; hh < kWgTH; hh += 2) {
                        // A: dY rows hh (j=0) and hh+1 (j=1): LBO = one tile row
                        const uint64_t da = make_smem_desc(ys + hh * (kWgTW * 128), kWgTW * 128, 1024);
                        const uint32_t acc = (kt > kt0 || hh > 0) ? 1u : 0u;
#pragma unroll
                        for (int v = 0; v < 2; ++v) {
                            // B: X local row hh + v, the three kw views 128 B apart
                            const uint64_t db = make_smem_desc(xs + (hh + v) * (kWgPitch * 128), 128, 1024);
                            umma_bf16(tmem_base + v * 192, da, db, idesc, acc);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kWg4Stages) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull_bar);
                aph ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int j = r >> 6, o = r & 63;
        uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int z = (int)(t / items), item = (int)(t % items);
            const int vp = item & 1, ob = (item >> 1) % p.oblocks, cc = (item >> 1) / p.oblocks;
            int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
            if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
            const bool have = kt1 > kt0;
            mbar_wait(tfull_bar, aph);
            tcgen05_fence_after();
            // split-K slice 2z + j; row o of the (O, 9C) gradient
            float* row = p.ws + (long long)(2 * z + j) * p.split_stride + (long long)(ob * 64 + o) * p.ldn + cc * 64;
#pragma unroll 1
            for (int v = 0; v < 2; ++v) {
                const int kh = 2 * vp + v - j;            // real filter row of this (view, j) half
#pragma unroll 1
                for (int kw = 0; kw < 3; ++kw) {
                    uint32_t a[32], b2[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + v * 192 + kw * 64, a);
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + v * 192 + kw * 64 + 32, b2);
                    tmem_wait_ld();
                    if (kh >= 0 && kh <= 2) {
                        float4* dst = reinterpret_cast<float4*>(row + (kh * 3 + kw) * p.C);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            dst[i] = have ? make_float4(__uint_as_float(a[4 * i]), __uint_as_float(a[4 * i + 1]),
                                                        __uint_as_float(a[4 * i + 2]), __uint_as_float(a[4 * i + 3]))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            dst[8 + i] = have ? make_float4(__uint_as_float(b2[4 * i]), __uint_as_float(b2[4 * i + 1]),
                                                            __uint_as_float(b2[4 * i + 2]), __uint_as_float(b2[4 * i + 3]))
                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            aph ^= 1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 forward / dgrad for NARROW outputs (64 or 128 channels), pixels on the N side of the MMA.
// Measured on B200 (tools/umma_rate_probe.cu): one tcgen05.mma M=128,K=16 with operands in shared memory never
// takes less than ~100 cycles (A-operand fetch), whatever N is; it only reaches the tensor peak for N >= 192.
// With pixels on M and N = O = 64 / 128 the generic kernel is therefore pinned at 32 % / 64 % of peak.  Here
// the roles are swapped:     D[row][n] = sum_{tap,c} Wm[row][tap*C+c] * X[pix(n) + s(tap)][c]
//   A (M = 128 rows)  = weight tile [128][64 c] per (tap, chunk), K-major, TMA box of the packed weight matrix
//   B (N = 256 rows)  = 256 pixels x 64 channels, one TMA box per (tap, chunk) at tap-shifted coordinates
// so every MMA is 128x256x16 (128 cycles, full rate).  Two row layouts:
//   P = 1  (O % 128 == 0):  row = output channel o (128 per m-block), 9 taps, weights = the ordinary packed filter.
//   P = 2  (O % 64 == 0):   row = (j, o), j = 0/1 selects output row 2i+j of a row pair, 64 channels per m-block.
//          Both halves read the SAME pixel tile, so the filter is re-expressed over 4 x 3 "virtual taps"
//          (input row 2i + vr - 1, vr = 0..3): rows j use W[kh = vr - j][kw] or zero ("pair-packed" weights,
//          unetca_pack_conv3x3_pair).  9 of the 12 MMA row blocks are useful -> 75 % of peak instead of 32 %.
//          The pixel tile is a lattice of even rows: a 5-D tensor map (C, W, parity, H/2, B).
// The accumulator comes out transposed (TMEM lane = channel, column = pixel): the epilogue writes it through a
// [pixel][64 ch] staging tile (2-byte stores, 64 contiguous bytes per warp: conflict free) and TMA-stores it; the
// BatchNorm partial sums are thread-local (one lane = one channel).
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) PixNParams {
    CUtensorMap mapX, mapW, mapOut;
    int tilesW, tilesI, nimg;
    int TW, TI, twShift, HP, W, P;      // HP = H / P rows of the (parity) lattice
    int ntaps, cchunks, num_m_blocks;
    signed char dw[12], par[12], off[12];
    float* stat_parts;
    int N;
    const float* ep_scale;       // optional per-channel affine + ReLU applied in the epilogue (eval-mode BatchNorm)
    const float* ep_shift;
    float* sq_parts;             // optional [nimg][gridDim.x][N]: per-image channel sums of the stored activations (SE squeeze)
};
constexpr int kPnStages = 3;
constexpr int kPnABytes = 128 * 128, kPnBBytes = 256 * 128, kPnStageBytes = kPnABytes + kPnBBytes;
constexpr int kPnOutBytes = 2 * 256 * 128;
constexpr int kPnStatBytes = 2 * 1024 * 4 + 256 * 4;
constexpr int kPnThreads = 320;          // producer, MMA issuer, eight epilogue warps
constexpr int kPnHandBytes = 2 * 128 * 4;                   // per-tile statistics handed from the second epilogue half to the first
constexpr int kPnSmemBytes = 1024 + kPnStages * kPnStageBytes + kPnOutBytes + kPnStatBytes + kPnHandBytes + 256;
static_assert(kPnSmemBytes <= 227 * 1024, "pixn conv: shared memory budget");

// write this CTA's per-image channel sums (the entries of sm_sq owned by epilogue thread r) to
// sq_parts[image][blockIdx.x][N] and clear them; P = 1: thread r owns channels mb*128 + r, P = 2: threads r < 64 own mb*64 + r
__device__ __forceinline__ void pixn_flush_sq(float* sq_parts, float* sm_sq, int image, int N, int P, int num_m_blocks, int r) {
    float* dst = sq_parts + ((long)image * gridDim.x + blockIdx.x) * N;
    if (P == 1) {
        for (int mb = 0; mb < num_m_blocks; ++mb) { dst[mb * 128 + r] = sm_sq[mb * 128 + r]; sm_sq[mb * 128 + r] = 0.f; }
    } else if (r < 64) {
        for (int mb = 0; mb < num_m_blocks; ++mb) { dst[mb * 64 + r] = sm_sq[mb * 64 + r]; sm_sq[mb * 64 + r] = 0.f; }
    }
}

// epilogue of the pixels-on-N kernel for one thread (= one accumulator row = one channel): 256 fp32 columns (pixels)
// -> bf16 -> staging[pixel][channel] (2-byte stores; a warp covers 64 contiguous bytes per pixel), plus the
// thread-local sum / sum of squares of the stored values.  PARTIAL masks pixels outside the image.
// BWD (haloed kernel only: 8-pixel-wide tiles): instead of sum / sum of squares, the ReLU + BatchNorm backward statistics
// of the tensor being written, s1 = sum dz, s2 = sum dz * (y - mean) with dz = value * (ba*y + bb > 0) and y the saved
// pre-BN conv output of the same pixel and channel (ybase -> this thread's channel at the tile's first pixel).
template <bool PARTIAL, bool BWD = false>
__device__ __forceinline__ void pixn_drain(uint32_t t_addr, uint32_t my_s, float& s1, float& s2, int x0, int i0,
                                           int twMask, int twShift, int W, int HP, bool act = false, float ea = 1.f,
                                           float eb = 0.f, const bf16* ybase = nullptr, long yrow = 0, int ypix = 0,
                                           float ba = 0.f, float bb = 0.f, float bm = 0.f, int cb0 = 0, int ncb = 8) {
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
#pragma unroll 1
    for (int cb = cb0; cb < cb0 + ncb; ++cb) {
        uint32_t v[32];
        tmem_ld32(t_addr + cb * 32, v);
        uint16_t yv[BWD ? 32 : 1];
        if (BWD) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int n = cb * 32 + i;
                const bool ok = !PARTIAL || ((x0 + (n & 7)) < W && (i0 + (n >> 3)) < HP);
                yv[BWD ? i : 0] = ok ? __ldg(reinterpret_cast<const unsigned short*>(ybase + (n >> 3) * yrow + (long)(n & 7) * ypix))
                                     : (uint16_t)0;
            }
        }
        tmem_wait_ld();
        const uint32_t base = my_s + cb * 32 * 128;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float f0 = __uint_as_float(v[i]), f1 = __uint_as_float(v[i + 1]);
            if (act) {           // inference: BatchNorm (running statistics folded into ea, eb) + ReLU in the epilogue
                f0 = fmaxf(fmaf(ea, f0, eb), 0.f);
                f1 = fmaxf(fmaf(ea, f1, eb), 0.f);
            }
            if (PARTIAL) {
                const int n = cb * 32 + i;
                if (!((x0 + (n & twMask)) < W && (i0 + (n >> twShift)) < HP)) f0 = 0.f;
                if (!((x0 + ((n + 1) & twMask)) < W && (i0 + ((n + 1) >> twShift)) < HP)) f1 = 0.f;
            }
            const uint32_t pk = pack_bf16x2(f0, f1);
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(base + i * 128), "h"((uint16_t)(pk & 0xffffu)) : "memory");
            asm volatile("st.shared.u16 [%0], %1;" ::"r"(base + (i + 1) * 128), "h"((uint16_t)(pk >> 16)) : "memory");
            const float r0 = __uint_as_float(pk << 16), r1 = __uint_as_float(pk & 0xffff0000u);
            if (BWD) {
                const float y0 = __uint_as_float((uint32_t)yv[BWD ? i : 0] << 16), y1 = __uint_as_float((uint32_t)yv[BWD ? i + 1 : 0] << 16);
                const float d0 = fmaf(ba, y0, bb) > 0.f ? r0 : 0.f, d1 = fmaf(ba, y1, bb) > 0.f ? r1 : 0.f;
                a1 += d0; a2 = fmaf(d0, y0 - bm, a2);
                b1 += d1; b2 = fmaf(d1, y1 - bm, b2);
            } else {
                a1 += r0; a2 = fmaf(r0, r0, a2);
                b1 += r1; b2 = fmaf(r1, r1, b2);
            }
        }
    }
    s1 = a1 + b1; s2 = a2 + b2;
}

// CL = 2: launched as clusters of two CTAs that walk the same (m-block, tap, chunk) sequence on neighbouring pixel
// tiles; each CTA fetches HALF of every weight tile and TMA-multicasts it into both shared memories, so the L2->SM
// traffic of the weight operand halves (this kernel is fill-bound: 48 KB per 512 MMA cycles without it).  A stage may
// only be refilled when BOTH CTAs' MMAs have drained it: the MMA commits are multicast to both empty barriers.
template <int CL>
__global__ void __launch_bounds__(kPnThreads, 1) tc_conv3x3_pixn_kernel(const __grid_constant__ PixNParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* out_stage = smem + kPnStages * kPnStageBytes;
    float* sm_stats = reinterpret_cast<float*>(out_stage + kPnOutBytes);
    float* sm_wpart = sm_stats + 2048;
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kPnOutBytes + kPnStatBytes + kPnHandBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kPnStages;
    uint64_t* tfull_bar = bars + 2 * kPnStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < kPnStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapW);
        prefetch_tmap(&p.mapOut);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (p.stat_parts || p.sq_parts) {
        for (int i = threadIdx.x; i < 2 * p.N; i += kPnThreads) sm_stats[i] = 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();              // the peer's barriers are initialised before anything is multicast
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // work: item u = (m-block, group of CL neighbouring pixel tiles); this CTA takes tile CL*group + rank.  A tile index
    // past the end is a phantom: its loads are zero-filled (batch coordinate out of range), its stores clipped.
    const int tiles_per_img = p.tilesW * p.tilesI;
    const long num_tiles = (long)tiles_per_img * p.nimg;
    const long num_work = ((num_tiles + CL - 1) / CL) * p.num_m_blocks;
    const int nk = p.ntaps * p.cchunks;
    const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
    const long first = blockIdx.x / CL, stride = gridDim.x / CL;
    constexpr uint16_t kAll = (uint16_t)((1u << CL) - 1);

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = first; t < num_work; t += stride) {
                const int mb = (int)(t % p.num_m_blocks);
                const long mt = (t / p.num_m_blocks) * CL + crank;
                const int tw = (int)(mt % p.tilesW), ti = (int)((mt / p.tilesW) % p.tilesI), b = (int)(mt / tiles_per_img);
                const int x0 = tw * p.TW, i0 = ti * p.TI;
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    for (int cc = 0; cc < p.cchunks; ++cc) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        uint8_t* sa = smem + s * kPnStageBytes;
                        mbar_expect_tx(&full_bar[s], kPnStageBytes);
                        if (CL == 1)
                            tma_load_4d(&p.mapW, &full_bar[s], sa, (tap * p.cchunks + cc) * 64, mb * 128, 0, 0);
                        else
                            tma_load_4d_mc(&p.mapW, &full_bar[s], sa + crank * (kPnABytes / CL), (tap * p.cchunks + cc) * 64,
                                           mb * 128 + crank * (128 / CL), 0, 0, kAll);
                        tma_load_5d(&p.mapX, &full_bar[s], sa + kPnABytes, cc * 64, x0 + p.dw[tap], p.par[tap],
                                    i0 + p.off[tap], b);
                        if (++s == kPnStages) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 0, 0);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (long t = first; t < num_work; t += stride) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kPnStageBytes);
                    const uint32_t sb = sa + kPnABytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                                  (kb | k) != 0);
                    if (CL == 1) umma_commit(&empty_bar[s]);
                    else umma_commit_mc(&empty_bar[s], kAll);
                    if (++s == kPnStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
                as ^= 1; if (as == 0) aph ^= 1;
            }
        }
    } else {
        // eight epilogue warps (2..9): warps w and w + 4 share a TMEM lane quarter and take 128 of the 256 pixel columns each
        // (the drain is a dependent chain one warp per scheduler issues at ~0.35 instructions per cycle; with a single
        // k-block per tile, as in the first convolution, it is the whole tile time); the second half hands its statistics
        // to the first through shared memory in a fixed order
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int r = q * 32 + lane;             // accumulator row = TMEM lane
        const int ep_tid = threadIdx.x - 64;     // 0..255
        const int box = r >> 6, oc = r & 63;
        uint8_t* my = out_stage + box * (256 * 128) + oc * 2;
        float* sm_hand = sm_wpart + 256;         // [2][128]: statistics of the second column half
        int as = 0; uint32_t aph = 0;
        int cur_b = -1;
        for (long t = first; t < num_work; t += stride) {
            const int mb = (int)(t % p.num_m_blocks);
            const long mt = (t / p.num_m_blocks) * CL + crank;
            const int tw = (int)(mt % p.tilesW), ti = (int)((mt / p.tilesW) % p.tilesI), b = (int)(mt / tiles_per_img);
            const int x0 = tw * p.TW, i0 = ti * p.TI;
            const bool full = (x0 + p.TW <= p.W) && (i0 + p.TI <= p.HP);
            mbar_wait(&tfull_bar[as], aph);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
            if (ep_tid == 0) tma_store_wait_read();          // staging drained by the previous tile's store
            named_bar_sync(1, 256);
            float s1 = 0.f, s2 = 0.f;
            const uint32_t my_s = smem_u32(my);
            const bool act = p.ep_scale != nullptr;
            const int ech = p.P == 1 ? mb * 128 + r : mb * 64 + (r & 63);
            const float ea = act ? __ldg(p.ep_scale + ech) : 1.f, eb = act ? __ldg(p.ep_shift + ech) : 0.f;
            if (full) pixn_drain<false>(t_addr, my_s, s1, s2, 0, 0, 0, 0, 0, 0, act, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, chalf * 4, 4);
            else pixn_drain<true>(t_addr, my_s, s1, s2, x0, i0, p.TW - 1, p.twShift, p.W, p.HP, act, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, chalf * 4, 4);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (chalf == 1) { sm_hand[r] = s1; sm_hand[128 + r] = s2; }
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (ep_tid == 0) {
                if (p.P == 1) {
                    tma_store_5d(&p.mapOut, out_stage, mb * 128, x0, 0, i0, b);
                    tma_store_5d(&p.mapOut, out_stage + 256 * 128, mb * 128 + 64, x0, 0, i0, b);
                } else {
                    tma_store_5d(&p.mapOut, out_stage, mb * 64, x0, 0, i0, b);
                    tma_store_5d(&p.mapOut, out_stage + 256 * 128, mb * 64, x0, 1, i0, b);
                }
                tma_store_commit();
            }
            // from here on only the first column half works (128 threads, named barrier 2); its partner's sums are in sm_hand
            if (chalf == 0) {
                s1 += sm_hand[r]; s2 += sm_hand[128 + r];
                if (p.stat_parts) {
                    if (p.P == 1) {
                        sm_stats[mb * 128 + r] += s1;
                        sm_stats[p.N + mb * 128 + r] += s2;
                    } else {
                        sm_wpart[r] = s1; sm_wpart[128 + r] = s2;
                        named_bar_sync(2, 128);
                        if (r < 64) {
                            sm_stats[mb * 64 + r] += sm_wpart[r] + sm_wpart[r + 64];
                            sm_stats[p.N + mb * 64 + r] += sm_wpart[128 + r] + sm_wpart[192 + r];
                        }
                    }
                }
                if (p.sq_parts && b < p.nimg) {
                    // SE squeeze of the stored activation: per-image channel sums, flushed when this CTA moves to the next image
                    if (b != cur_b) {
                        if (cur_b >= 0) pixn_flush_sq(p.sq_parts, sm_stats, cur_b, p.N, p.P, p.num_m_blocks, r);
                        cur_b = b;
                    }
                    if (p.P == 1) sm_stats[mb * 128 + r] += s1;
                    else {
                        sm_wpart[r] = s1;
                        named_bar_sync(2, 128);
                        if (r < 64) sm_stats[mb * 64 + r] += sm_wpart[r] + sm_wpart[r + 64];
                        named_bar_sync(2, 128);
                    }
                }
            }
            as ^= 1; if (as == 0) aph ^= 1;
        }
        if (chalf == 0 && p.sq_parts && cur_b >= 0) pixn_flush_sq(p.sq_parts, sm_stats, cur_b, p.N, p.P, p.num_m_blocks, r);
        if (ep_tid == 0) tma_store_wait_all();
        named_bar_sync(1, 256);
        if (p.stat_parts) {
            float* dst = p.stat_parts + (long)blockIdx.x * 2 * p.N;
            for (int i = ep_tid; i < 2 * p.N; i += 256) dst[i] = sm_stats[i];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();              // the peer may still multicast into / arrive on this CTA's shared memory
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// DIRECT-STORE epilogue shared by the haloed and the resident-filter row-pair kernels.
// Epilogue of one thread (= one accumulator row: a channel, or (j, channel) in the row-pair layout) over the 128 columns [cb0*32, cb0*32 + 128): fp32 -> (BN + ReLU)
// -> bf16 -> global memory (a warp = 32 channels of one pixel = 64 contiguous bytes), plus either sum / sum of squares of the
// stored values or (BWD) the ReLU + BatchNorm backward statistics against the saved conv output of the same pixels.
// Column n = (pair row i = n >> 3, x = n & 7).  FULL: the whole tile lies inside the image (no bounds checks).
// LDO / LDY: compile-time pixel strides of the output / saved tensor (0 = take the run-time value) — with the two layouts the
// model uses (dense 64, channel half of a 128-wide concat buffer) every store address is row pointer + immediate.
template <bool FULL, bool BWD, int LDO, int LDY, bool ACT = false, bool BIAS = false>
__device__ __forceinline__ void rp64_drain(uint32_t t_addr, int cb0, bf16* __restrict__ obase, long row_stride, int ldo_rt, int ni,
                                           int nx, bool act_rt, float ea, float eb, const bf16* __restrict__ ybase, long yrow_stride,
                                           int ldy_rt, float ba, float bb, float bm, float& s1, float& s2) {
    const int ldo = LDO ? LDO : ldo_rt, ldy = LDY ? LDY : ldy_rt;
    const bool act = ACT || (!FULL && act_rt);          // the inference epilogue: specialised on full tiles, run-time flag on partial ones
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
    unsigned short* prow = reinterpret_cast<unsigned short*>(obase) + (long)cb0 * 4 * row_stride;
    const unsigned short* yrow = reinterpret_cast<const unsigned short*>(ybase) + (BWD ? (long)cb0 * 4 * yrow_stride : 0);
#pragma unroll 1
    for (int cb = cb0; cb < cb0 + 4; ++cb) {
        uint32_t v[32];
        tmem_ld32(t_addr + cb * 32, v);
        unsigned short yv[BWD ? 32 : 1];
        if (BWD) {
            // the saved conv output of the same 32 pixels: all loads in flight before the accumulator wait
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const bool ok = FULL || ((cb * 4 + ii) < ni && x < nx);
                    yv[BWD ? ii * 8 + x : 0] = ok ? __ldg(yrow + ii * yrow_stride + x * ldy) : (unsigned short)0;
                }
            }
            yrow += 4 * yrow_stride;
        }
        tmem_wait_ld();
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const bool rok = FULL || (cb * 4 + ii) < ni;
#pragma unroll
            for (int x = 0; x < 8; x += 2) {
                float f0 = __uint_as_float(v[ii * 8 + x]), f1 = __uint_as_float(v[ii * 8 + x + 1]);
                if (act) {
                    f0 = fmaxf(fmaf(ea, f0, eb), 0.f);
                    f1 = fmaxf(fmaf(ea, f1, eb), 0.f);
                }
                if (BIAS) { f0 += eb; f1 += eb; }
                const bool ok0 = rok && (FULL || x < nx), ok1 = rok && (FULL || x + 1 < nx);
                if (!FULL) { if (!ok0) f0 = 0.f; if (!ok1) f1 = 0.f; }
                const uint32_t pk = pack_bf16x2(f0, f1);
                if (ok0) prow[x * ldo] = (unsigned short)(pk & 0xffffu);
                if (ok1) prow[(x + 1) * ldo] = (unsigned short)(pk >> 16);
                const float r0 = __uint_as_float(pk << 16), r1 = __uint_as_float(pk & 0xffff0000u);
                if (BWD) {
                    // dz = value where the forward ReLU was open; (a2, b2) collect sum dz*y — the mean is taken out once at the end
                    const float y0 = __uint_as_float((uint32_t)yv[BWD ? ii * 8 + x : 0] << 16);
                    const float y1 = __uint_as_float((uint32_t)yv[BWD ? ii * 8 + x + 1 : 0] << 16);
                    if (fmaf(ba, y0, bb) > 0.f) { a1 += r0; a2 = fmaf(r0, y0, a2); }
                    if (fmaf(ba, y1, bb) > 0.f) { b1 += r1; b2 = fmaf(r1, y1, b2); }
                } else {
                    a1 += r0; a2 = fmaf(r0, r0, a2);
                    b1 += r1; b2 = fmaf(r1, r1, b2);
                }
            }
            prow += row_stride;
        }
    }
    s1 = a1 + b1;
    s2 = BWD ? fmaf(-bm, s1, a2 + b2) : a2 + b2;        // sum dz*(y - mean) = sum dz*y - mean * sum dz
}

// The fused BatchNorm-backward statistics need the saved conv output y of every pixel the thread writes.  Those are 2-byte
// loads that miss L1 (each tile is read once): issued next to their use they put ~1 us of L2 / HBM latency into every
// 32-column chunk.  So they run one chunk AHEAD: the first chunk's loads are issued before the thread waits for the tile's
// accumulator (they do not depend on it), each further chunk's loads before the previous chunk is processed.
// (two bf16 values per register: pixels x and x + 1 of a row)
template <bool FULL, int LDY>
__device__ __forceinline__ void bwd_load_y(uint32_t (&yv)[16], const unsigned short* __restrict__ yrow, long yrow_stride,
                                           int ldy_rt, int cb, int ni, int nx) {
    const int ldy = LDY ? LDY : ldy_rt;
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
        for (int x = 0; x < 8; x += 2) {
            const bool rok = FULL || (cb * 4 + ii) < ni;
            const uint32_t lo = (rok && (FULL || x < nx)) ? (uint32_t)__ldg(yrow + ii * yrow_stride + x * ldy) : 0u;
            const uint32_t hi = (rok && (FULL || x + 1 < nx)) ? (uint32_t)__ldg(yrow + ii * yrow_stride + (x + 1) * ldy) : 0u;
            yv[ii * 4 + (x >> 1)] = lo | (hi << 16);
        }
    }
}
// yv: the saved conv output of all 128 pixels of this thread's column half (bwd_load_all_y), loaded by the caller BEFORE
// it waited for the accumulator — the loads depend on nothing the MMA produces and are complete by the time the tile is
template <bool FULL, int LDY>
__device__ __forceinline__ void bwd_load_all_y(uint32_t (&yv)[4][16], const bf16* __restrict__ ybase, long yrow_stride,
                                               int ldy_rt, int cb0, int ni, int nx) {
    const unsigned short* yrow = reinterpret_cast<const unsigned short*>(ybase) + (long)cb0 * 4 * yrow_stride;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        bwd_load_y<FULL, LDY>(yv[k], yrow, yrow_stride, ldy_rt, cb0 + k, ni, nx);
        yrow += 4 * yrow_stride;
    }
}
template <bool FULL, int LDO>
__device__ __forceinline__ void rp64_drain_bwd(uint32_t t_addr, int cb0, bf16* __restrict__ obase, long row_stride, int ldo_rt,
                                               int ni, int nx, float ba, float bb, float bm, uint32_t (&yv)[4][16], float& s1,
                                               float& s2) {
    const int ldo = LDO ? LDO : ldo_rt;
    float a1 = 0.f, a2 = 0.f, b1 = 0.f, b2 = 0.f;
    unsigned short* prow = reinterpret_cast<unsigned short*>(obase) + (long)cb0 * 4 * row_stride;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int cb = cb0 + k;
        uint32_t v[32];
        tmem_ld32(t_addr + cb * 32, v);
        tmem_wait_ld();
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const bool rok = FULL || (cb * 4 + ii) < ni;
#pragma unroll
            for (int x = 0; x < 8; x += 2) {
                float f0 = __uint_as_float(v[ii * 8 + x]), f1 = __uint_as_float(v[ii * 8 + x + 1]);
                const bool ok0 = rok && (FULL || x < nx), ok1 = rok && (FULL || x + 1 < nx);
                if (!FULL) { if (!ok0) f0 = 0.f; if (!ok1) f1 = 0.f; }
                const uint32_t pk = pack_bf16x2(f0, f1);
                if (ok0) prow[x * ldo] = (unsigned short)(pk & 0xffffu);
                if (ok1) prow[(x + 1) * ldo] = (unsigned short)(pk >> 16);
                const float r0 = __uint_as_float(pk << 16), r1 = __uint_as_float(pk & 0xffff0000u);
                // dz = value where the forward ReLU was open; (a2, b2) collect sum dz*y — the mean is taken out once at the end
                const uint32_t yp = yv[k][ii * 4 + (x >> 1)];
                const float y0 = __uint_as_float(yp << 16), y1 = __uint_as_float(yp & 0xffff0000u);
                if (fmaf(ba, y0, bb) > 0.f) { a1 += r0; a2 = fmaf(r0, y0, a2); }
                if (fmaf(ba, y1, bb) > 0.f) { b1 += r1; b2 = fmaf(r1, y1, b2); }
            }
            prow += row_stride;
        }
    }
    s1 = a1 + b1;
    s2 = fmaf(-bm, s1, a2 + b2);                         // sum dz*(y - mean) = sum dz*y - mean * sum dz
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 forward / dgrad, pixels on N (as above, 128 output channels per m-block) WITH tap reuse from one haloed
// activation tile.  Why: a 128x256x16 MMA reads 12 KB of operands from shared memory in 128 cycles (96 B/cycle);
// the one-box-per-tap kernels additionally WRITE 48 KB per 512 cycles of TMA fill into the same shared memory
// (94 B/cycle), and the two together exceed what the SM's shared memory sustains — every such kernel here, and
// cuBLAS's single-CTA tiles, level off near 70 % of the MMA rate.  Here the pixel tile is 8 wide x 32 high, its
// 10 x 34 haloed input tile (43.5 KB per 64-channel chunk) is loaded ONCE, and the nine taps are K-major B
// descriptors into it (start shifted by (kh*10 + kw) pixel rows, SBO = 10 rows = 1280 B; the swizzle is a function
// of the absolute address, tools/umma_shift_probe.cu).  Fill drops from 432 KB to 188 KB per chunk (41 B/cycle).
// Warps: 0 = activation producer, 1 = MMA issuer, 2..5 = epilogue (shared with the kernel above), 6 = weight producer.
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) HpixParams {
    CUtensorMap mapX, mapW;
    CUtensorMap mapX2;           // second activation source (channels >= xsplit), see tma_load_x2
    int xsplit;
    bf16* out;
    int ldo;
    // optional second destination: output channels >= split go to out2 (pixel stride ldo2, channel - split) — the dgrad of a
    // decoder block's first conv writes d(skip) and d(upsampled) as two dense tensors instead of one [.., 2C] buffer whose
    // halves every consumer would then read as 128-byte rows at a 256-byte stride
    bf16* out2;
    int ldo2, split;
    int tilesW, tilesH, nimg, H, W;
    int cchunks, num_m_blocks;
    float* stat_parts;
    int N;
    const float* ep_scale;
    const float* ep_shift;
    float* sq_parts;             // optional [nimg][gridDim.x][N] per-image channel sums (SE squeeze of the activation)
    // optional fused ReLU + BatchNorm backward statistics of the tensor being written (see pixn_drain<.., BWD>); the sums
    // go to stat_parts in place of sum / sum of squares
    const bf16* bwd_y;
    int bwd_ldy;
    const float* bwd_scale;
    const float* bwd_shift;
    const float* bwd_mean;
};
constexpr int kHpThreads = 352;          // X producer, MMA issuer, eight epilogue warps, weight producer
constexpr int kHpTW = 8, kHpTH = 32, kHpPitch = kHpTW + 2;
constexpr int kHpXBox = kHpPitch * (kHpTH + 2) * 128;      // 43520 bytes written by TMA
constexpr int kHpXBytes = 44 * 1024;                       // ring slot
constexpr int kHpXS = 3, kHpWS = 5;                        // (the 64 KB output staging tile of the TMA-store epilogue is gone)
constexpr int kHpWBytes = 128 * 128;
constexpr int kHpSmemBytes = 1024 + kHpXS * kHpXBytes + kHpWS * kHpWBytes + kPnStatBytes + 256;
static_assert(kHpSmemBytes <= 227 * 1024, "hpix conv: shared memory budget");

__global__ void __launch_bounds__(kHpThreads, 1) tc_conv3x3_hpix_kernel(const __grid_constant__ HpixParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* x_ring = smem;
    uint8_t* w_ring = smem + kHpXS * kHpXBytes;
    float* sm_stats = reinterpret_cast<float*>(w_ring + kHpWS * kHpWBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(w_ring + kHpWS * kHpWBytes + kPnStatBytes);
    uint64_t* xfull = bars;
    uint64_t* xempty = bars + kHpXS;
    uint64_t* wfull = bars + 2 * kHpXS;
    uint64_t* wempty = wfull + kHpWS;
    uint64_t* tfull_bar = wempty + kHpWS;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < kHpXS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
        for (int i = 0; i < kHpWS; ++i) { mbar_init(&wfull[i], 1); mbar_init(&wempty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapW);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (p.stat_parts || p.sq_parts) {
        for (int i = threadIdx.x; i < 2 * p.N; i += kHpThreads) sm_stats[i] = 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const long num_work = (long)tiles_per_img * p.nimg * p.num_m_blocks;

    if (warp == 0) {
        // ================================ activation producer: one haloed tile per (item, chunk) ==============
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int mt = (int)(t / p.num_m_blocks);
                const int tw = mt % p.tilesW, th = (mt / p.tilesW) % p.tilesH, b = mt / tiles_per_img;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    mbar_wait(&xempty[s], ph ^ 1);
                    mbar_expect_tx(&xfull[s], kHpXBox);
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &xfull[s], x_ring + s * kHpXBytes, cc * 64, tw * kHpTW - 1, th * kHpTH - 1, b);
                    if (++s == kHpXS) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 10) {
        // ================================ weight producer: one [128 rows][64] box per (chunk, tap) ============
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int mb = (int)(t % p.num_m_blocks);
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&wempty[s], ph ^ 1);
                        mbar_expect_tx(&wfull[s], kHpWBytes);
                        tma_load_4d(&p.mapW, &wfull[s], w_ring + s * kHpWBytes, (tap * p.cchunks + cc) * 64, mb * 128, 0, 0);
                        if (++s == kHpWS) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 0, 0);
            int xs = 0; uint32_t xph = 0;
            int ws = 0; uint32_t wph = 0;
            int as = 0; uint32_t aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    mbar_wait(&xfull[xs], xph);
                    const uint32_t xa = smem_u32(x_ring + xs * kHpXBytes);
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&wfull[ws], wph);
                        tcgen05_fence_after();
                        const uint32_t sa = smem_u32(w_ring + ws * kHpWBytes);
                        const uint32_t sb = xa + ((tap / 3) * kHpPitch + (tap % 3)) * 128;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024),
                                      make_smem_desc(sb + k * 32, 16, kHpPitch * 128), idesc, (cc | tap | k) != 0);
                        umma_commit(&wempty[ws]);
                        if (++ws == kHpWS) { ws = 0; wph ^= 1; }
                    }
                    umma_commit(&xempty[xs]);
                    if (++xs == kHpXS) { xs = 0; xph ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
                as ^= 1; if (as == 0) aph ^= 1;
            }
        }
    } else {
        // ================================ epilogue (warps 2..9): lane = channel, column = pixel ================
        // Eight warps (w and w + 4 share a TMEM lane quarter and take 128 of the 256 pixel columns each) that write their
        // bf16 values STRAIGHT to global memory — a warp = 32 channels of one pixel = 64 contiguous bytes, two full
        // sectors.  The staging tile + TMA store this replaces cost ~1500 shared-memory wavefronts per tile (1024 two-byte
        // stores + the store's reads) on the pipe that also feeds the MMA: 22 % of the pipe at 128 input channels, where
        // operands (75 %) + fill (33 %) already exceed it; its 64 KB now hold a third activation stage and a fifth
        // weight stage.  The second column half hands its statistics to the first through shared memory (fixed order).
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const int ep_tid = threadIdx.x - 64;      // 0..255
        float* sm_wpart = sm_stats + 2048;        // [2][128]
        const bool act = p.ep_scale != nullptr;
        const bool bwd = p.bwd_y != nullptr;
        const long yrow_stride = (long)p.W * p.bwd_ldy;
        int as = 0; uint32_t aph = 0;
        int cur_b = -1;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int mb = (int)(t % p.num_m_blocks);
            const int mt = (int)(t / p.num_m_blocks);
            const int tw = mt % p.tilesW, th = (mt / p.tilesW) % p.tilesH, b = mt / tiles_per_img;
            const int x0 = tw * kHpTW, h0 = th * kHpTH;
            const int nx = p.W - x0 < kHpTW ? p.W - x0 : kHpTW, ni = p.H - h0 < kHpTH ? p.H - h0 : kHpTH;
            const bool full = nx == kHpTW && ni == kHpTH;
            const int ch = mb * 128 + r;
            const long pix0 = ((long)b * p.H + h0) * p.W + x0;
            const bool second = p.out2 != nullptr && ch >= p.split;
            const int ldo = second ? p.ldo2 : p.ldo;
            bf16* obase = second ? p.out2 + pix0 * ldo + (ch - p.split) : p.out + pix0 * ldo + ch;
            const long row_stride = (long)p.W * ldo;
            const bf16* ybase = bwd ? p.bwd_y + pix0 * p.bwd_ldy + ch : nullptr;
            uint32_t yv[4][16];
            if (bwd) {          // the saved-output loads go out before the accumulator is waited for
                if (full) bwd_load_all_y<true, 0>(yv, ybase, yrow_stride, p.bwd_ldy, chalf * 4, ni, nx);
                else bwd_load_all_y<false, 0>(yv, ybase, yrow_stride, p.bwd_ldy, chalf * 4, ni, nx);
            }
            mbar_wait(&tfull_bar[as], aph);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
            float s1 = 0.f, s2 = 0.f;
            if (bwd) {
                const float ba = __ldg(p.bwd_scale + ch), bb = __ldg(p.bwd_shift + ch), bm = __ldg(p.bwd_mean + ch);
                if (full) rp64_drain_bwd<true, 0>(t_addr, chalf * 4, obase, row_stride, ldo, ni, nx, ba, bb, bm, yv, s1, s2);
                else rp64_drain_bwd<false, 0>(t_addr, chalf * 4, obase, row_stride, ldo, ni, nx, ba, bb, bm, yv, s1, s2);
            } else {
                const float ea = act ? __ldg(p.ep_scale + ch) : 1.f, eb = act ? __ldg(p.ep_shift + ch) : 0.f;
                if (full && act) rp64_drain<true, false, 0, 0, true>(t_addr, chalf * 4, obase, row_stride, ldo, ni, nx, true, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
                else if (full && ldo == 128) rp64_drain<true, false, 128, 0>(t_addr, chalf * 4, obase, row_stride, 128, ni, nx, false, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
                else if (full && ldo == 256) rp64_drain<true, false, 256, 0>(t_addr, chalf * 4, obase, row_stride, 256, ni, nx, false, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
                else if (full && ldo == 512) rp64_drain<true, false, 512, 0>(t_addr, chalf * 4, obase, row_stride, 512, ni, nx, false, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
                else if (full) rp64_drain<true, false, 0, 0>(t_addr, chalf * 4, obase, row_stride, ldo, ni, nx, false, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
                else rp64_drain<false, false, 0, 0>(t_addr, chalf * 4, obase, row_stride, ldo, ni, nx, act, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (p.stat_parts || p.sq_parts) {
                if (chalf == 1) { sm_wpart[r] = s1; sm_wpart[128 + r] = s2; }
                named_bar_sync(1, 256);
                if (chalf == 0) {
                    if (p.stat_parts) {
                        sm_stats[mb * 128 + r] += s1 + sm_wpart[r];
                        sm_stats[p.N + mb * 128 + r] += s2 + sm_wpart[128 + r];
                    }
                    if (p.sq_parts) {
                        if (b != cur_b) {
                            if (cur_b >= 0) pixn_flush_sq(p.sq_parts, sm_stats, cur_b, p.N, 1, p.num_m_blocks, r);
                            cur_b = b;
                        }
                        sm_stats[mb * 128 + r] += s1 + sm_wpart[r];
                    }
                }
                named_bar_sync(1, 256);         // sm_wpart may be rewritten
            }
            as ^= 1; if (as == 0) aph ^= 1;
        }
        if (chalf == 0 && p.sq_parts && cur_b >= 0) pixn_flush_sq(p.sq_parts, sm_stats, cur_b, p.N, 1, p.num_m_blocks, r);
        named_bar_sync(1, 256);
        if (p.stat_parts) {
            float* dst = p.stat_parts + (long)blockIdx.x * 2 * p.N;
            for (int i = ep_tid; i < 2 * p.N; i += 256) dst[i] = sm_stats[i];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// ConvTranspose2d(k = 2, s = 2) forward (UCA:133..136 up1..up4), sub-pixel channels on M, input pixels on N:
//   D[(d, e, o)][(i, x)] = sum_c Wt[(d, e, o)][c] * X[h0 + i][x0 + x][c],      out[b][2(h0+i) + d][2(x0+x) + e][o] = D + bias[o]
// One m-block = 128 of the 4*Cout rows, one pixel tile = 8 x 32 input pixels (N = 256), K = Cin in 64-channel chunks, each one
// [256 px][64] activation box + one [128][64] weight box.  The generic pixels-on-M kernel this replaces is EPILOGUE-bound on
// this op — K is short (2..16 chunks) and every accumulator element becomes an output element, which its four epilogue warps
// drain through a staging tile and four strided TMA stores per tile, one tile at a time; at 128 -> 64 channels it ran at 0.29
// of the tensor rate and 0.53 of the op's HBM floor.  Here the eight epilogue warps of the haloed conv kernel write straight to
// global memory (a warp = 32 channels of one output pixel = 64 contiguous bytes; the sub-pixel (d, e) is a per-thread constant
// folded into the base pointer, the output pixel pitch is 2*ldo) while the next tile's MMAs fill the other accumulator buffer.
// Work order: m-blocks of one pixel tile are adjacent work items, so the CTAs of a wave re-read the activation tile from L2.
// Warps: 0 = producer, 1 = MMA issuer, 2..9 = epilogue.
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) CtfParams {
    CUtensorMap mapX, mapW;
    bf16* out;
    int ldo;
    const float* bias;
    int tilesW, tilesH, nimg, h, w;      // input extents
    int cchunks, num_m_blocks, Cout;
};
constexpr int kCtThreads = 320;
constexpr int kCtXBytes = 256 * 128, kCtWBytes = 128 * 128, kCtStageBytes = kCtXBytes + kCtWBytes, kCtStages = 4;
constexpr int kCtSmemBytes = 1024 + kCtStages * kCtStageBytes + 256;

__global__ void __launch_bounds__(kCtThreads, 1) tc_convT_fwd_kernel(const __grid_constant__ CtfParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCtStages * kCtStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kCtStages;
    uint64_t* tfull_bar = bars + 2 * kCtStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kCtStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapW);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const long num_work = (long)tiles_per_img * p.nimg * p.num_m_blocks;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int mb = (int)(t % p.num_m_blocks);
                const int mt = (int)(t / p.num_m_blocks);
                const int tw = mt % p.tilesW, th = (mt / p.tilesW) % p.tilesH, b = mt / tiles_per_img;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + s * kCtStageBytes;
                    mbar_expect_tx(&full_bar[s], kCtStageBytes);
                    tma_load_4d(&p.mapX, &full_bar[s], st, cc * 64, tw * kHpTW, th * kHpTH, b);
                    tma_load_4d(&p.mapW, &full_bar[s], st + kCtXBytes, cc * 64, mb * 128, 0, 0);
                    if (++s == kCtStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 0, 0);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t sb = smem_u32(smem + s * kCtStageBytes);
                    const uint32_t sa = sb + kCtXBytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                                  (cc | k) != 0);
                    umma_commit(&empty_bar[s]);
                    if (++s == kCtStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
                as ^= 1; if (as == 0) aph ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const long W2 = 2L * p.w;
        const int ldo2 = 2 * p.ldo;
        const long row_stride = 2 * W2 * p.ldo;
        int as = 0; uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int mb = (int)(t % p.num_m_blocks);
            const int mt = (int)(t / p.num_m_blocks);
            const int tw = mt % p.tilesW, th = (mt / p.tilesW) % p.tilesH, b = mt / tiles_per_img;
            const int x0 = tw * kHpTW, h0 = th * kHpTH;
            const int nx = p.w - x0 < kHpTW ? p.w - x0 : kHpTW, ni = p.h - h0 < kHpTH ? p.h - h0 : kHpTH;
            const bool full = nx == kHpTW && ni == kHpTH;
            const int ch = mb * 128 + r;
            const int de = ch / p.Cout, o = ch - de * p.Cout;
            bf16* obase = p.out + (((long)b * 2 * p.h + 2 * h0 + (de >> 1)) * W2 + 2 * x0 + (de & 1)) * p.ldo + o;
            const float eb = __ldg(p.bias + o);
            mbar_wait(&tfull_bar[as], aph);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
            float s1, s2;
            if (full && ldo2 == 256) rp64_drain<true, false, 256, 0, false, true>(t_addr, chalf * 4, obase, row_stride, 256, ni, nx, false, 1.f, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
            else if (full && ldo2 == 512) rp64_drain<true, false, 512, 0, false, true>(t_addr, chalf * 4, obase, row_stride, 512, ni, nx, false, 1.f, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
            else if (full) rp64_drain<true, false, 0, 0, false, true>(t_addr, chalf * 4, obase, row_stride, ldo2, ni, nx, false, 1.f, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
            else rp64_drain<false, false, 0, 0, false, true>(t_addr, chalf * 4, obase, row_stride, ldo2, ni, nx, false, 1.f, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, s1, s2);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            as ^= 1; if (as == 0) aph ^= 1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// ConvTranspose2d(k = 2, s = 2) weight gradient with a 256 x 256 output tile per work item        (UCA:345 through up1..up4)
//   dW[c][(d, e, o)] = sum_{b,i,j} X[b,i,j,c] * dOut[b, 2i+d, 2j+e, o]           (K = input pixels, both operands MN-major)
// The generic split-K kernel (128 x 256 tile) is bound by L2 -> shared-memory fill on this op: nothing is reused across MMAs,
// every 128x256x16 MMA needs 12 KB of fresh operands, and halving the tile width (BLOCK_N = 128) takes it from 0.47 to 0.71 ms
// at 1024 -> 512 — time follows the bytes filled.  Here one k-tile of 64 pixels brings four [64 px][64 c] activation boxes and
// four [64 px][64] gradient boxes (64 KB) for EIGHT MMAs into two 128 x 256 accumulators (all 512 TMEM columns): 8 KB per MMA.
// The accumulator is single-buffered; an item's mainloop is > 100 k-tiles, its epilogue (256 KB of fp32 partials) ~3 %.
// ---------------------------------------------------------------------------------------------------------
constexpr int kCwStages = 3;
constexpr int kCwABytes = 4 * kBoxBytesWg, kCwBBytes = 4 * kBoxBytesWg, kCwStageBytes = kCwABytes + kCwBBytes;
constexpr int kCwSmemBytes = 1024 + kCwStages * kCwStageBytes + 256;

__global__ void __launch_bounds__(kTcThreads, 1) tc_convT_wgrad_kernel(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kCwStages * kCwStageBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kCwStages;
    uint64_t* tfull_bar = bars + 2 * kCwStages;
    uint64_t* tempty_bar = tfull_bar + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 1);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kCwStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(tfull_bar, 1); mbar_init(tempty_bar, 4);
        fence_barrier_init();
        prefetch_tmap(&p.mapA[0]);
        for (int i = 0; i < 4; ++i) prefetch_tmap(&p.mapB[i]);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesH;
    const long mn_items = (long)p.num_m_blocks * p.num_n_blocks;
    const long num_work = mn_items * p.nsplit;
    const int kt_per_split = (p.ktiles_total + p.nsplit - 1) / p.nsplit;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / mn_items);
                const int nb = (int)((t % mn_items) % p.num_n_blocks);
                const int mb = (int)((t % mn_items) / p.num_n_blocks);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                for (int kt = kt0; kt < kt1; ++kt) {
                    const int tw = kt % p.tilesW, th = (kt / p.tilesW) % p.tilesH, b = kt / tiles_per_img;
                    const int w0 = tw * p.TW, h0 = th * p.TH;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * kCwStageBytes;
                    mbar_expect_tx(&full_bar[s], kCwStageBytes);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        tma_load_4d(&p.mapA[0], &full_bar[s], sa + i * kBoxBytesWg, (mb * 4 + i) * 64, w0, h0, b);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int q = nb * 4 + j;
                        tma_load_4d(&p.mapB[q / p.b_chunks_per_map], &full_bar[s], sa + kCwABytes + j * kBoxBytesWg,
                                    (q % p.b_chunks_per_map) * 64, w0, h0, b);
                    }
                    if (++s == kCwStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 1, 1);
            int s = 0; uint32_t ph = 0;
            uint32_t aph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int z = (int)(t / mn_items);
                int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
                if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
                const int nk = kt1 - kt0;
                mbar_wait(tempty_bar, aph ^ 1);
                tcgen05_fence_after();
                for (int kb = 0; kb < nk; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kCwStageBytes);
                    const uint32_t sb = sa + kCwABytes;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t db = make_smem_desc(sb + k * 2048, kBoxBytesWg, 1024);
                        umma_bf16(tmem_base, make_smem_desc(sa + k * 2048, kBoxBytesWg, 1024), db, idesc, (kb | k) != 0);
                        umma_bf16(tmem_base + 256, make_smem_desc(sa + 2 * kBoxBytesWg + k * 2048, kBoxBytesWg, 1024), db, idesc,
                                  (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                    if (++s == kCwStages) { s = 0; ph ^= 1; }
                }
                umma_commit(tfull_bar);
                aph ^= 1;
            }
        }
    } else {
        const int q = warp & 3;
        const int r = q * 32 + lane;
        uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int z = (int)(t / mn_items);
            const int nb = (int)((t % mn_items) % p.num_n_blocks);
            const int mb = (int)((t % mn_items) / p.num_n_blocks);
            int kt0 = z * kt_per_split, kt1 = kt0 + kt_per_split;
            if (kt1 > p.ktiles_total) kt1 = p.ktiles_total;
            const bool have = kt1 > kt0;
            float* wsz = p.ws + (long long)z * p.split_stride;
            mbar_wait(tfull_bar, aph);
            tcgen05_fence_after();
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                const int m = mb * 256 + h * 128 + r;
                const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + h * 256;
#pragma unroll 1
                for (int c32 = 0; c32 < 8; ++c32) {
                    uint32_t v[32];
                    tmem_ld32(t_addr + c32 * 32, v);
                    tmem_wait_ld();
                    if (m < p.m_valid) {
                        float4* dst = reinterpret_cast<float4*>(wsz + (long long)m * p.ldn + nb * 256 + c32 * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            dst[i] = have ? make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                        __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar);
            aph ^= 1;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 forward / dgrad for 64 -> 64 channels (the two full-resolution DoubleConv layers, 4 launches per step):
// the row-pair layout of tc_conv3x3_pixn_kernel<P = 2> with the tap reuse of the haloed kernel and a RESIDENT filter.
//   D[(j, o)][(i, x)] = sum_{vr, kw, c} Wv[(j, o)][vr][kw][c] * X[2(i0+i) + vr - 1][x0 + x + kw - 1][c]
// M = 128 = 2 output rows of a row pair x 64 channels, N = 256 = 32 pair rows x 8 columns, 4 x 3 virtual taps of
// which rows j use W[kh = vr - j] or zero (9 of 12 row blocks useful).  What bounded the per-tap kernel was shared-
// memory fill: 12 x (16 KB weights + 32 KB pixels) per 48 MMAs = 94 B/cycle next to 96 B/cycle of operand reads.  Here:
//   * the input rows of a tile are two haloed LATTICE tiles (even image rows 2i: vr = 1, 3; odd rows 2i+1: vr = 0, 2),
//     each [33 lattice rows][10 px][64 ch] = 42 KB through the 5-D (C, W, parity, H/2, B) tensor map; the six taps of a
//     lattice are K-major B descriptors into it, start shifted by (dv*10 + kw) pixel rows, SBO = 10 px = 1280 B (8 pixels
//     of one pair row per 8-row group) — 84 KB of fill per tile instead of 384 KB;
//   * the whole filter stays in shared memory as thirteen 8 KB blocks [64 o][64 c]:  Z b2 b1 b0 Z b2 b1 b0 Z b2 b1 b0 Z
//     (one run per kw, b_kh = W[.][kh][kw][.], Z = zeros).  The A operand of virtual tap (vr, kw) is the 16 KB starting
//     at block (run + 2 - vr): rows j = 0 read b_vr, rows j = 1 the next block b_(vr-1), and the out-of-range filter
//     rows fall on a Z block — no pair-packed copy of the filter exists any more (TMA boxes of the ordinary [O][9*C]);
//   * the epilogue writes its bf16 values straight to global memory (a warp = 32 channels of one pixel = 64 contiguous
//     bytes, two full sectors): no staging tile, so the ~1500 wavefronts per tile of staging stores and TMA-store reads
//     leave the shared-memory pipe to the MMA (operands 4608 + fill 660 wavefronts per 6144 MMA cycles).
// Warps: 0 = producer, 1 = MMA issuer, 2..5 = epilogue.  ~190 KB of shared memory.
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) Rp64Params {
    CUtensorMap mapX, mapW;
    bf16* out;
    int ldo;
    int tilesW, tilesI, nimg, HP, W;
    float* stat_parts;           // optional [gridDim.x][2][64]: sum / sum of squares of the stored values
    const float* ep_scale;       // optional eval-mode BatchNorm + ReLU in the epilogue
    const float* ep_shift;
    float* sq_parts;             // optional [nimg][gridDim.x][64]: per-image channel sums of the stored activations (SE squeeze)
    // optional fused ReLU + BatchNorm backward statistics of the tensor being written (dA1 = this dgrad's output):
    // sum dz, sum dz * (y - mean) per channel with dz = dA1 * (a*y + b > 0), y = bwd_y (the saved pre-BN conv output)
    const bf16* bwd_y;
    int bwd_ldy;
    const float* bwd_scale;
    const float* bwd_shift;
    const float* bwd_mean;
};
constexpr int kRpTW = 8, kRpTI = 32, kRpPitch = kRpTW + 2, kRpRows = kRpTI + 1;
constexpr int kRpXBox = kRpPitch * kRpRows * 128;           // 42240 bytes written by TMA per lattice tile
constexpr int kRpXBytes = 42 * 1024;                        // slot
constexpr int kRpWBlock = 64 * 128;                         // [64 o][64 c] bf16
constexpr int kRpWBytes = 13 * kRpWBlock;
constexpr int kRpStatBytes = (2 * 64 + 2 * 256 + 64) * 4;
constexpr int kRpSmemBytes = 1024 + kRpWBytes + 2 * kRpXBytes + kRpStatBytes + 256;
static_assert(kRpSmemBytes <= 227 * 1024, "rp64 conv: shared memory budget");

constexpr int kRpThreads = 320;          // producer, MMA issuer, eight epilogue warps

__global__ void __launch_bounds__(kRpThreads, 1) tc_conv3x3_rp64_kernel(const __grid_constant__ Rp64Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* w_res = smem;
    uint8_t* x_ring = smem + kRpWBytes;
    float* sm_stats = reinterpret_cast<float*>(x_ring + 2 * kRpXBytes);      // [2][64]
    float* sm_wpart = sm_stats + 128;                                         // [2][256]
    float* sm_sq = sm_wpart + 512;                                            // [64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(x_ring + 2 * kRpXBytes + kRpStatBytes);
    uint64_t* xfull = bars;            // [2]
    uint64_t* xempty = bars + 2;       // [2]
    uint64_t* wfull = bars + 4;
    uint64_t* tfull_bar = bars + 5;    // [2]
    uint64_t* tempty_bar = bars + 7;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    // warp index through a shuffle: tells the compiler that the role branches below are warp-uniform, so the epilogue's global
    // accesses keep their descriptors in uniform registers instead of re-deriving them per access (3 R2UR per store otherwise)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
        mbar_init(wfull, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapW);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    // the four zero blocks of the resident filter (blocks 0, 4, 8, 12) and the statistics
    for (int i = threadIdx.x; i < 4 * (kRpWBlock / 16); i += kRpThreads) {
        const int blk = (i / (kRpWBlock / 16)) * 4, off = (i % (kRpWBlock / 16)) * 16;
        *reinterpret_cast<uint4*>(w_res + blk * kRpWBlock + off) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = threadIdx.x; i < kRpStatBytes / 4; i += kRpThreads) sm_stats[i] = 0.f;
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesW * p.tilesI;
    const long num_work = (long)tiles_per_img * p.nimg;

    if (warp == 0) {
        if (lane == 0) {
            // the filter, once: block (kh, kw) of the packed [64][9*64] matrix -> slot 1 + 4*kw + (2 - kh)
            mbar_expect_tx(wfull, 9 * kRpWBlock);
            for (int kh = 0; kh < 3; ++kh)
                for (int kw = 0; kw < 3; ++kw)
                    tma_load_4d(&p.mapW, wfull, w_res + (1 + 4 * kw + (2 - kh)) * kRpWBlock, (kh * 3 + kw) * 64, 0, 0, 0);
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int tw = (int)(t % p.tilesW), ti = (int)((t / p.tilesW) % p.tilesI), b = (int)(t / tiles_per_img);
                const int x0 = tw * kRpTW, i0 = ti * kRpTI;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    // half 0: even lattice rows i0 .. i0+32 (vr = 1, 3); half 1: odd lattice rows i0-1 .. i0+31 (vr = 0, 2)
                    mbar_wait(&xempty[s], ph ^ 1);
                    mbar_expect_tx(&xfull[s], kRpXBox);
                    tma_load_5d(&p.mapX, &xfull[s], x_ring + s * kRpXBytes, 0, x0 - 1, half, i0 - half, b);
                    if (++s == 2) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 256, 0, 0);
            mbar_wait(wfull, 0);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            const uint32_t wb = smem_u32(w_res);
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    mbar_wait(&xfull[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xa = smem_u32(x_ring + s * kRpXBytes);
#pragma unroll
                    for (int dv = 0; dv < 2; ++dv) {
                        const int vr = half == 0 ? 1 + 2 * dv : 2 * dv;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const uint32_t sa = wb + (uint32_t)(1 + 4 * kw + 2 - vr) * kRpWBlock;
                            const uint32_t sb = xa + (uint32_t)(dv * kRpPitch + kw) * 128;
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024),
                                          make_smem_desc(sb + k * 32, 16, kRpPitch * 128), idesc, (half | dv | kw | k) != 0);
                        }
                    }
                    umma_commit(&xempty[s]);
                    if (++s == 2) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
                as ^= 1; if (as == 0) aph ^= 1;
            }
        }
    } else {
        // ================================ epilogue (warps 2..9) =======================================================
        // TMEM lane = (j, channel); warps w and w + 4 share a lane quarter and take 128 of the 256 columns (pair row, x) each
        const int q = warp & 3;
        const int chalf = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const int ep_tid = threadIdx.x - 64;     // 0..255
        const int j = r >> 6, oc = r & 63;
        const bool act = p.ep_scale != nullptr;
        const float ea = act ? __ldg(p.ep_scale + oc) : 1.f, eb = act ? __ldg(p.ep_shift + oc) : 0.f;
        const bool bwd = p.bwd_y != nullptr;
        const float ba = bwd ? __ldg(p.bwd_scale + oc) : 0.f, bb = bwd ? __ldg(p.bwd_shift + oc) : 0.f;
        const float bm = bwd ? __ldg(p.bwd_mean + oc) : 0.f;
        const long row_stride = 2L * p.W * p.ldo;                             // one pair row down
        const long yrow_stride = 2L * p.W * p.bwd_ldy;
        int as = 0; uint32_t aph = 0;
        int cur_b = -1;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            const int tw = (int)(t % p.tilesW), ti = (int)((t / p.tilesW) % p.tilesI), b = (int)(t / tiles_per_img);
            const int x0 = tw * kRpTW, i0 = ti * kRpTI;
            // element (i, x) of this thread's channel: out[((b*H + 2*(i0+i) + j) * W + x0 + x) * ldo + oc]
            const long pix0 = ((long)b * 2 * p.HP + 2 * i0 + j) * p.W + x0;
            bf16* obase = p.out + pix0 * p.ldo + oc;
            const bf16* ybase = bwd ? p.bwd_y + pix0 * p.bwd_ldy + oc : nullptr;
            const int nx = p.W - x0 < kRpTW ? p.W - x0 : kRpTW;              // valid columns
            const int ni = p.HP - i0 < kRpTI ? p.HP - i0 : kRpTI;            // valid pair rows
            const bool full = nx == kRpTW && ni == kRpTI;
            const bool dense = p.ldo == 64 && p.bwd_ldy == 64;
            uint32_t yv[4][16];
            if (bwd) {          // the saved-output loads go out before the accumulator is waited for
                if (full && dense) bwd_load_all_y<true, 64>(yv, ybase, yrow_stride, 64, chalf * 4, ni, nx);
                else if (full) bwd_load_all_y<true, 0>(yv, ybase, yrow_stride, p.bwd_ldy, chalf * 4, ni, nx);
                else bwd_load_all_y<false, 0>(yv, ybase, yrow_stride, p.bwd_ldy, chalf * 4, ni, nx);
            }
            mbar_wait(&tfull_bar[as], aph);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
            float a1, a2;
            if (bwd) {
                if (full && dense) rp64_drain_bwd<true, 64>(t_addr, chalf * 4, obase, row_stride, 64, ni, nx, ba, bb, bm, yv, a1, a2);
                else if (full) rp64_drain_bwd<true, 0>(t_addr, chalf * 4, obase, row_stride, p.ldo, ni, nx, ba, bb, bm, yv, a1, a2);
                else rp64_drain_bwd<false, 0>(t_addr, chalf * 4, obase, row_stride, p.ldo, ni, nx, ba, bb, bm, yv, a1, a2);
            } else {
                if (full && act) rp64_drain<true, false, 0, 0, true>(t_addr, chalf * 4, obase, row_stride, p.ldo, ni, nx, true, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, a1, a2);
                else if (full && p.ldo == 64) rp64_drain<true, false, 64, 0>(t_addr, chalf * 4, obase, row_stride, 64, ni, nx, false, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, a1, a2);
                else if (full && p.ldo == 128) rp64_drain<true, false, 128, 0>(t_addr, chalf * 4, obase, row_stride, 128, ni, nx, false, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, a1, a2);
                else if (full) rp64_drain<true, false, 0, 0>(t_addr, chalf * 4, obase, row_stride, p.ldo, ni, nx, false, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, a1, a2);
                else rp64_drain<false, false, 0, 0>(t_addr, chalf * 4, obase, row_stride, p.ldo, ni, nx, act, ea, eb, nullptr, 0, 0, 0.f, 0.f, 0.f, a1, a2);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            if (p.stat_parts) {
                sm_wpart[chalf * 128 + r] = a1; sm_wpart[256 + chalf * 128 + r] = a2;
                named_bar_sync(1, 256);
                if (ep_tid < 64) {
                    sm_stats[ep_tid] += (sm_wpart[ep_tid] + sm_wpart[ep_tid + 64]) + (sm_wpart[128 + ep_tid] + sm_wpart[192 + ep_tid]);
                    sm_stats[64 + ep_tid] += (sm_wpart[256 + ep_tid] + sm_wpart[320 + ep_tid]) + (sm_wpart[384 + ep_tid] + sm_wpart[448 + ep_tid]);
                }
                named_bar_sync(1, 256);
            }
            if (p.sq_parts) {
                // SE squeeze of the stored activation: per-image channel sums, flushed when this CTA moves to the next image
                if (b != cur_b) {
                    if (cur_b >= 0 && ep_tid < 64) {
                        p.sq_parts[((long)cur_b * gridDim.x + blockIdx.x) * 64 + ep_tid] = sm_sq[ep_tid];
                        sm_sq[ep_tid] = 0.f;
                    }
                    cur_b = b;
                }
                sm_wpart[chalf * 128 + r] = a1;
                named_bar_sync(1, 256);
                if (ep_tid < 64) sm_sq[ep_tid] += (sm_wpart[ep_tid] + sm_wpart[ep_tid + 64]) + (sm_wpart[128 + ep_tid] + sm_wpart[192 + ep_tid]);
                named_bar_sync(1, 256);
            }
            as ^= 1; if (as == 0) aph ^= 1;
        }
        if (p.sq_parts && cur_b >= 0 && ep_tid < 64) p.sq_parts[((long)cur_b * gridDim.x + blockIdx.x) * 64 + ep_tid] = sm_sq[ep_tid];
        named_bar_sync(1, 256);
        if (p.stat_parts) {
            float* dst = p.stat_parts + (long)blockIdx.x * 128;
            if (ep_tid < 128) dst[ep_tid] = sm_stats[ep_tid];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// conv3x3 forward / dgrad for 64 output channels, the three kw taps on the MMA's N side ("kw-stacked").
// Pixels stay on M (128 = 4 rows x 32 columns of the image), N = 192 = 3 kw taps x 64 output channels:
//   D[p][(kw, o)] = sum_{kh, c} X[p + (kh-1) rows][c] * W[o][kh][kw][c]          (only VERTICAL shifts of the A view)
//   out[(y, x)][o] = D[(y, x-1)][(0,o)] + D[(y, x)][(1,o)] + D[(y, x+1)][(2,o)]   (horizontal shift-add in the epilogue)
// One tcgen05.mma 128x192x16 costs its ~100-cycle A-fetch floor for 96 cycles of work (N = 64 alone: 32), there are no
// wasted rows (the row-pair layout spends 25 % of its MMAs on zeros) and the A operand needs no horizontal halo: the
// kh views are the same dense [6 rows][32 px] tile at row offsets.  The whole filter of the layer (3 x C/64 boxes of
// [192][64], 72 / 144 KB) stays resident in shared memory, so the only fill traffic is 24 KB of activations per
// 12 MMAs.  The neighbour terms come from warp shuffles (a warp = one 32-pixel image row of the tile); the two edge
// columns of a tile are halo, so a tile yields 30 x 4 outputs.
// ---------------------------------------------------------------------------------------------------------
struct alignas(64) KwParams {
    CUtensorMap mapX, mapW, mapOut;
    CUtensorMap mapX2;           // second activation source (channels >= xsplit), see tma_load_x2
    int xsplit;
    int tilesX, tilesY, nimg, H, W;
    int cchunks;
    float* stat_parts;
    const float* ep_scale;
    const float* ep_shift;
};
constexpr int kKwTW = 32, kKwTH = 4, kKwOutW = kKwTW - 2;
constexpr int kKwXBytes = kKwTW * (kKwTH + 2) * 128;      // 24 KB: [6 rows][32 px][64 ch]
constexpr int kKwWBox = 192 * 128;                        // 24 KB: [(kw,o)][64 c]
constexpr int kKwOutTile = kKwTH * (kKwTW - 2) * 128;     // staging tile: [4 x 30 px][64 ch] = 15 KB
constexpr int kKwOutBytes = 2 * kKwOutTile;               // double-buffered: the TMA store of tile i drains under tile i+1
constexpr int kKwStatBytes = 2 * 64 * 4 + 16 * 64 * 4;
constexpr int kKwThreads = 320;                          // producer, MMA issuer, eight epilogue warps
static int kw_smem_bytes(int cchunks, int xs) { return 3 * cchunks * kKwWBox + xs * kKwXBytes + kKwOutBytes + kKwStatBytes + 256; }

template <int XS>
__global__ void __launch_bounds__(kKwThreads, 1) tc_conv3x3_kw_kernel(const __grid_constant__ KwParams p) {
    // the budget has no room for an alignment pad (resident filter + ring + staging fill the 227 KB): the array is
    // declared 1024-byte aligned and the kernel traps if the driver ever places it otherwise
    extern __shared__ __align__(1024) uint8_t kw_smem[];
    uint8_t* smem = kw_smem;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* w_res = smem;                                         // resident filter: [kh][cc] boxes
    uint8_t* x_ring = w_res + 3 * p.cchunks * kKwWBox;
    uint8_t* out_stage = x_ring + XS * kKwXBytes;
    float* sm_stats = reinterpret_cast<float*>(out_stage + kKwOutBytes);      // [2][64]
    float* sm_wpart = sm_stats + 128;                                         // [8 warps][2][64]
    uint64_t* bars = reinterpret_cast<uint64_t*>(out_stage + kKwOutBytes + kKwStatBytes);
    uint64_t* xfull = bars;
    uint64_t* xempty = bars + XS;
    uint64_t* wfull = bars + 2 * XS;
    uint64_t* tfull_bar = wfull + 1;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler

    if (threadIdx.x == 0) {
        for (int i = 0; i < XS; ++i) { mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1); }
        mbar_init(wfull, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
        fence_barrier_init();
        prefetch_tmap(&p.mapX);
        prefetch_tmap(&p.mapW);
        prefetch_tmap(&p.mapOut);
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (p.stat_parts) {
        for (int i = threadIdx.x; i < 128; i += kKwThreads) sm_stats[i] = 0.f;
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tiles_per_img = p.tilesX * p.tilesY;
    const long num_work = (long)tiles_per_img * p.nimg;

    if (warp == 0) {
        if (lane == 0) {
            // the whole filter once
            mbar_expect_tx(wfull, 3 * p.cchunks * kKwWBox);
            for (int i = 0; i < 3 * p.cchunks; ++i) tma_load_4d(&p.mapW, wfull, w_res + i * kKwWBox, 0, i * 192, 0, 0);
            int s = 0; uint32_t ph = 0;
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                const int tx = (int)(t % p.tilesX), ty = (int)((t / p.tilesX) % p.tilesY), b = (int)(t / tiles_per_img);
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    mbar_wait(&xempty[s], ph ^ 1);
                    mbar_expect_tx(&xfull[s], kKwXBytes);
                    tma_load_x2(&p.mapX, &p.mapX2, p.xsplit, &xfull[s], x_ring + s * kKwXBytes, cc * 64, tx * kKwOutW - 1, ty * kKwTH - 1, b);
                    if (++s == XS) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, 192, 0, 0);
            mbar_wait(wfull, 0);
            int s = 0; uint32_t ph = 0;
            int as = 0; uint32_t aph = 0;
            const uint32_t wb = smem_u32(w_res);
            for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int cc = 0; cc < p.cchunks; ++cc) {
                    mbar_wait(&xfull[s], ph);
                    tcgen05_fence_after();
                    const uint32_t xa = smem_u32(x_ring + s * kKwXBytes);
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const uint32_t sa = xa + kh * (kKwTW * 128);
                        const uint32_t sb = wb + (kh * p.cchunks + cc) * kKwWBox;
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                                      (cc | kh | k) != 0);
                    }
                    umma_commit(&xempty[s]);
                    if (++s == XS) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
                as ^= 1; if (as == 0) aph ^= 1;
            }
        }
    } else {
        // ================================ epilogue: lane = pixel column of the tile, warp pair = tile row =====
        // Eight warps: one warp per scheduler issues ~0.35 instructions per cycle through this dependent shuffle / add /
        // pack chain and the 24 MMAs of a tile finished in half the time the drain took (tensor pipe 47 % active);
        // warps w and w + 4 share a TMEM lane quarter and take 32 of the 64 output channels each.
        const int q = warp & 3;                   // TMEM lane quarter == tile row
        const int half = (warp - 2) >> 2;         // which 32 output channels
        const int ep_tid = threadIdx.x - 64;
        int as = 0; uint32_t aph = 0;
        for (long t = blockIdx.x; t < num_work; t += gridDim.x) {
            uint8_t* stage = out_stage + as * kKwOutTile;    // staging buffer of this tile (alternates like the accumulator)
            const int tx = (int)(t % p.tilesX), ty = (int)((t / p.tilesX) % p.tilesY), b = (int)(t / tiles_per_img);
            const int x0 = tx * kKwOutW, y0 = ty * kKwTH;
            const int xo = x0 + lane - 1, yo = y0 + q;       // the output pixel this thread assembles (lanes 1..30)
            const bool inner = lane >= 1 && lane <= kKwOutW;
            const bool valid = inner && xo < p.W && yo < p.H;
            mbar_wait(&tfull_bar[as], aph);
            tcgen05_fence_after();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + as * 256;
            if (ep_tid == 0) tma_store_wait_read1();         // the store issued two tiles ago has drained this buffer
            named_bar_sync(1, 256);
            const int nrow = q * kKwOutW + lane - 1;         // staging row of this thread's pixel
            uint8_t* row = stage + nrow * 128;
            {
                uint32_t a[32], c1[32], c2[32];
                tmem_ld32(t_addr + 0 * 64 + half * 32, a);
                tmem_ld32(t_addr + 1 * 64 + half * 32, c1);
                tmem_ld32(t_addr + 2 * 64 + half * 32, c2);
                tmem_wait_ld();
                float f[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(a[i]), 1);      // kw = 0 from column x-1
                    const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(c2[i]), 1);  // kw = 2 from column x+1
                    float t = (left + __uint_as_float(c1[i])) + right;
                    if (p.ep_scale) t = fmaxf(fmaf(__ldg(p.ep_scale + half * 32 + i), t, __ldg(p.ep_shift + half * 32 + i)), 0.f);
                    f[i] = valid ? t : 0.f;
                }
                if (inner) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint4 pk;
                        pk.x = pack_bf16x2(f[c * 8 + 0], f[c * 8 + 1]); pk.y = pack_bf16x2(f[c * 8 + 2], f[c * 8 + 3]);
                        pk.z = pack_bf16x2(f[c * 8 + 4], f[c * 8 + 5]); pk.w = pack_bf16x2(f[c * 8 + 6], f[c * 8 + 7]);
                        const int chunk = half * 4 + c;
                        *reinterpret_cast<uint4*>(row + ((chunk ^ (nrow & 7)) << 4)) = pk;
                    }
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
            fence_proxy_async_smem();
            named_bar_sync(1, 256);
            if (ep_tid == 0) {
                tma_store_4d(&p.mapOut, stage, 0, x0, y0, b);
                tma_store_commit();
            }
            if (p.stat_parts) {
                // per-channel sum / sum of squares of the staged bf16 tile (120 rows x 8 chunks of 8 channels)
                const int ch = ep_tid & 7, r0 = ep_tid >> 3;
                float s1[8], s2[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
                for (int rr = r0; rr < kKwTH * kKwOutW; rr += 32) {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
                    const uint32_t w4[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float lo = __uint_as_float(w4[i] << 16), hi = __uint_as_float(w4[i] & 0xffff0000u);
                        s1[2 * i] += lo; s2[2 * i] = fmaf(lo, lo, s2[2 * i]);
                        s1[2 * i + 1] += hi; s2[2 * i + 1] = fmaf(hi, hi, s2[2 * i + 1]);
                    }
                }
#pragma unroll
                for (int off = 8; off < 32; off <<= 1) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], off);
                        s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], off);
                    }
                }
                const int ew = ep_tid >> 5;
                if (lane < 8) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        sm_wpart[(ew * 2 + 0) * 64 + ch * 8 + i] = s1[i];
                        sm_wpart[(ew * 2 + 1) * 64 + ch * 8 + i] = s2[i];
                    }
                }
                named_bar_sync(1, 256);
                if (ep_tid < 128) {
                    const int which = ep_tid >> 6, col = ep_tid & 63;
                    float tsum = 0.f;
#pragma unroll
                    for (int w8i = 0; w8i < 8; ++w8i) tsum += sm_wpart[(w8i * 2 + which) * 64 + col];
                    sm_stats[which * 64 + col] += tsum;
                }
            }
            as ^= 1; if (as == 0) aph ^= 1;
        }
        if (ep_tid == 0) tma_store_wait_all();
        named_bar_sync(1, 256);
        if (p.stat_parts) {
            float* dst = p.stat_parts + (long)blockIdx.x * 128;
            for (int i = ep_tid; i < 128; i += 256) dst[i] = sm_stats[i];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// kw-stacked filter: dst[((kh*cch + cc)*3 + kw)*64 + o][c_l] = src[o][(kh*3 + kw)*C + cc*64 + c_l]   (64 output rows)
__global__ void pack_kw_kernel(const bf16* __restrict__ src, int ld, bf16* __restrict__ dst, int C) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cch = C / 64;
    if (i >= 9L * cch * 64 * 64) return;
    const int cl = (int)(i & 63), o = (int)((i >> 6) & 63);
    const int blk = (int)(i >> 12);                       // (kh*cch + cc)*3 + kw
    const int kw = blk % 3, cc = (blk / 3) % cch, kh = blk / (3 * cch);
    dst[i] = src[(long)o * ld + (kh * 3 + kw) * C + cc * 64 + cl];
}

// pair-packed weights for the P = 2 layout: dst[(g*128 + j*64 + o)][(vr*3+kw)*C + c] = src[(g*64+o)][((vr-j)*3+kw)*C + c]
// when 0 <= vr-j <= 2, else 0.  src: K-major packed filter [rows][ld] (k = tap*C + c), rows a multiple of 64.
__global__ void pack_pair_kernel(const bf16* __restrict__ src, int ld, bf16* __restrict__ dst, int rows, int C) {
    const long total = (long)2 * rows * 12 * C;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int k = (int)(i % (12 * C));
    const int row = (int)(i / (12 * C));
    const int g = row / 128, j = (row / 64) & 1, o = row & 63;
    const int v = k / C, c = k % C;
    const int vr = v / 3, kw = v % 3;
    const int kh = vr - j;
    dst[i] = (kh >= 0 && kh <= 2) ? src[(long)(g * 64 + o) * ld + (kh * 3 + kw) * C + c] : __float2bfloat16_rn(0.f);
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps, tiling, launch
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 4-D bf16 map: dims (C, W, H, B) with element strides (1, sw, sh, sb); box (64, bw, bh, 1); SWIZZLE_128B.
static int make_map(CUtensorMap* m, const void* base, long C, long W, long H, long B, long sw, long sh, long sb,
                    int bw, int bh) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return UNETCA_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        set_error("tensor map: base/strides must be 16-byte aligned");
        return UNETCA_ERR_ARG;
    }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims %ld %ld %ld %ld strides %ld %ld %ld box %d %d", (int)r, C, W,
                  H, B, sw, sh, sb, bw, bh);
        return UNETCA_ERR_CUDA;
    }
    return 0;
}

// pick a (TW, TH) pixel box with TW*TH == npx minimising padded work
static void pick_tile(int H, int W, int npx, int* TW, int* TH) {
    long best = -1;
    for (int tw = 8; tw <= npx && tw <= 256; tw *= 2) {
        const int th = npx / tw;
        if (th > 256) continue;
        const long work = (long)ceil_div(W, tw) * tw * ceil_div(H, th) * th;
        if (best < 0 || work < best || (work == best && tw > *TW)) { best = work; *TW = tw; *TH = th; }
    }
}

template <int BLOCK_N, bool WGRAD>
static int launch_tc(const TcParams& p, long num_work, cudaStream_t st, const char* what) {
    using Cfg = TcCfg<BLOCK_N, WGRAD>;
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_kernel<BLOCK_N, WGRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES);
        if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    long grid = num_work < num_sms() ? num_work : num_sms();
    if (grid < 1) grid = 1;
    tc_kernel<BLOCK_N, WGRAD><<<(int)grid, kTcThreads, Cfg::SMEM_BYTES, st>>>(p);
    int rc = check_launch(what);
    return rc < 0 ? rc : (int)grid;
}

template <bool WGRAD>
static int launch_tc_n(int block_n, const TcParams& p, long num_work, cudaStream_t st, const char* what) {
    switch (block_n) {
        case 64: return launch_tc<64, WGRAD>(p, num_work, st, what);
        case 128: return launch_tc<128, WGRAD>(p, num_work, st, what);
        case 256: return launch_tc<256, WGRAD>(p, num_work, st, what);
    }
    set_error("%s: BLOCK_N %d unsupported", what, block_n);
    return UNETCA_ERR_ARG;
}

static int g_force_block_n = 0;
static int g_wgrad_narrow = 0;
static int pick_block_n(int N) {
    if (g_force_block_n && N % g_force_block_n == 0) return g_force_block_n;
    if (N % 256 == 0) return 256;    // 128x256 tiles: 96 B/cycle/SM of operand fill instead of 128 (L2->SM is the limit)
    if (N % 128 == 0) return 128;
    return 64;
}

static int g_no_halo = 0;
static int g_wgrad_waves = 6;      // conv3x3 weight gradients: work items per SM that the split-K factor aims for
static int g_convT_wide = 1;       // ConvTranspose forward / wgrad: N blocks of 256 across the four sub-pixel maps

// make_map variant without swizzle (linear [pixel][64 ch] staging tiles of the pixels-on-N epilogue)
static int make_map_linear(CUtensorMap* m, const void* base, long C, long W, long H, long B, long sw, long sh, long sb,
                           int bw, int bh) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return UNETCA_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sw * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (((uintptr_t)base & 15) || (strides[0] & 15)) { set_error("tensor map: base/strides must be 16-byte aligned"); return UNETCA_ERR_ARG; }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (linear) failed (%d)", (int)r); return UNETCA_ERR_CUDA; }
    return 0;
}

// Plain (unswizzled) 4-D NHWC map for the streamed elementwise kernels: dims (C, W, H, B), pixel stride ld elements,
// box (C, bw, bh, 1); elem_bytes 2 (bf16) or 1 (uint8).  Out-of-bounds loads are zero-filled, stores clipped.
// (external linkage: used by elementwise.cu)
int make_tmap_nhwc(CUtensorMap* m, const void* base, int elem_bytes, long C, long W, long H, long B, long ld, int bw, int bh) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return UNETCA_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)ld * elem_bytes, (cuuint64_t)W * ld * elem_bytes, (cuuint64_t)H * W * ld * elem_bytes};
    cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (((uintptr_t)base & 15) || (strides[0] & 15) || C > 256 || bw > 256 || bh > 256 || ((C * elem_bytes) & 15)) {
        set_error("tensor map (nhwc): base/strides must be 16-byte aligned, box dims <= 256"); return UNETCA_ERR_ARG;
    }
    CUresult r = enc(m, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (nhwc) failed (%d)", (int)r); return UNETCA_ERR_CUDA; }
    return 0;
}

// optional second activation source of a conv whose input is a channel concat kept as two dense tensors
struct XSrc2 { const void* x2; int ldx2; int xsplit; };

// fused ReLU + BatchNorm backward statistics of a dgrad's output (saved pre-BN conv output y + the BN constants)
struct BwdStats { const void* y; int ldy; const float* scale; const float* shift; const float* mean; };

// conv3x3 forward / dgrad for O % 128 == 0 through the haloed pixels-on-N kernel; w = packed filter [O][9*C]
static int launch_hpix(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W, int C, int O,
                       float* stat_parts, cudaStream_t st, const float* ep_scale = nullptr, const float* ep_shift = nullptr,
                       float* sq_parts = nullptr, const BwdStats* bwd = nullptr, void* y2 = nullptr, int ldy2 = 0, int split = 0,
                       const XSrc2* xs2 = nullptr) {
    HpixParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    const int C1 = xs2 ? xs2->xsplit : C;
    if ((rc = make_map(&p.mapX, x, C1, W, H, B, ldx, (long)W * ldx, (long)H * W * ldx, kHpPitch, kHpTH + 2)) < 0) return rc;
    if (xs2) {
        if ((rc = make_map(&p.mapX2, xs2->x2, C - C1, W, H, B, xs2->ldx2, (long)W * xs2->ldx2, (long)H * W * xs2->ldx2, kHpPitch,
                           kHpTH + 2)) < 0) return rc;
        p.xsplit = C1;
    }
    if ((rc = make_map(&p.mapW, w, 9L * C, O, 1, 1, ldk, (long)O * ldk, (long)O * ldk, 128, 1)) < 0) return rc;
    p.out = (bf16*)y; p.ldo = ldy;
    p.out2 = (bf16*)y2; p.ldo2 = ldy2; p.split = split;
    p.tilesW = ceil_div(W, kHpTW); p.tilesH = ceil_div(H, kHpTH); p.nimg = B; p.H = H; p.W = W;
    p.cchunks = C / 64; p.num_m_blocks = O / 128;
    p.stat_parts = stat_parts; p.N = O;
    p.ep_scale = ep_scale; p.ep_shift = ep_shift; p.sq_parts = sq_parts;
    if (bwd) {
        p.bwd_y = (const bf16*)bwd->y; p.bwd_ldy = bwd->ldy;
        p.bwd_scale = bwd->scale; p.bwd_shift = bwd->shift; p.bwd_mean = bwd->mean;
    }
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_hpix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kHpSmemBytes);
        if (e != cudaSuccess) { set_error("tc_conv3x3 (hpix): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    const long num_work = (long)p.tilesW * p.tilesH * B * p.num_m_blocks;
    long grid = num_work < num_sms() ? num_work : num_sms();
    if (grid < 1) grid = 1;
    tc_conv3x3_hpix_kernel<<<(int)grid, kHpThreads, kHpSmemBytes, st>>>(p);
    rc = check_launch("tc_conv3x3_fwd (hpix)");
    return rc < 0 ? rc : (int)grid;
}

static int g_no_kw = 0;

// conv3x3 forward / dgrad for exactly 64 output channels and C in {64, 128}: kw-stacked kernel; w_kw from unetca_tc_pack_kw
static int launch_kw(const void* x, int ldx, const void* w_kw, void* y, int ldy, int B, int H, int W, int C,
                     float* stat_parts, cudaStream_t st, const float* ep_scale = nullptr, const float* ep_shift = nullptr,
                     const XSrc2* xs2 = nullptr) {
    KwParams p;
    memset(&p, 0, sizeof(p));
    const int cch = C / 64;
    int rc;
    const int C1 = xs2 ? xs2->xsplit : C;
    if ((rc = make_map(&p.mapX, x, C1, W, H, B, ldx, (long)W * ldx, (long)H * W * ldx, kKwTW, kKwTH + 2)) < 0) return rc;
    if (xs2) {
        if ((rc = make_map(&p.mapX2, xs2->x2, C - C1, W, H, B, xs2->ldx2, (long)W * xs2->ldx2, (long)H * W * xs2->ldx2, kKwTW,
                           kKwTH + 2)) < 0) return rc;
        p.xsplit = C1;
    }
    const long rows = 9L * cch * 64;
    if ((rc = make_map(&p.mapW, w_kw, 64, rows, 1, 1, 64, rows * 64, rows * 64, 192, 1)) < 0) return rc;
    if ((rc = make_map(&p.mapOut, y, 64, W, H, B, ldy, (long)W * ldy, (long)H * W * ldy, kKwOutW, kKwTH)) < 0) return rc;
    p.tilesX = ceil_div(W, kKwOutW); p.tilesY = ceil_div(H, kKwTH); p.nimg = B; p.H = H; p.W = W;
    p.cchunks = cch;
    p.stat_parts = stat_parts;
    p.ep_scale = ep_scale; p.ep_shift = ep_shift;
    const int xs = cch == 1 ? 4 : 2;
    const int smem = kw_smem_bytes(cch, xs);
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_kw_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kw_smem_bytes(1, 4));
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc_conv3x3_kw_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kw_smem_bytes(2, 2));
        if (e != cudaSuccess) { set_error("tc_conv3x3 (kw): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    const long num_work = (long)p.tilesX * p.tilesY * B;
    long grid = num_work < num_sms() ? num_work : num_sms();
    if (grid < 1) grid = 1;
    if (cch == 1) tc_conv3x3_kw_kernel<4><<<(int)grid, kKwThreads, smem, st>>>(p);
    else tc_conv3x3_kw_kernel<2><<<(int)grid, kKwThreads, smem, st>>>(p);
    rc = check_launch("tc_conv3x3_fwd (kw)");
    return rc < 0 ? rc : (int)grid;
}

// 5-D bf16 map over an NHWC activation seen as (C, W, P, H/P, B): P = 2 splits the rows into an even and an odd
// lattice (row = 2*i + parity).  box (64, bw, 1, bi, 1).
static int make_map5(CUtensorMap* m, const void* base, long C, long W, long H, long B, long ld, int P, int bw, int bi,
                     bool swizzle) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled not available"); return UNETCA_ERR_CUDA; }
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)P, (cuuint64_t)(H / P), (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)P * W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)bw, 1, (cuuint32_t)bi, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (((uintptr_t)base & 15) || (strides[0] & 15)) { set_error("tensor map: base/strides must be 16-byte aligned"); return UNETCA_ERR_ARG; }
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (5-D) failed (%d): dims %ld %ld %d %ld %ld box %d %d", (int)r, C, W, P, H / P, B, bw, bi);
        return UNETCA_ERR_CUDA;
    }
    return 0;
}

static int g_no_pixn = 0;
static int g_pixn_cluster = 2;     // CTAs per cluster sharing the weight operand by TMA multicast (1 or 2)

static int launch_pixn_kernel(const PixNParams& p, int CL, cudaStream_t st, const char* what) {
    const long num_tiles = (long)p.tilesW * p.tilesI * p.nimg;
    const long num_work = ((num_tiles + CL - 1) / CL) * p.num_m_blocks;          // items per cluster-walk
    long grid = num_work * CL < num_sms() ? num_work * CL : (num_sms() / CL) * CL;
    if (grid < CL) grid = CL;
    if (CL == 1) {
        tc_conv3x3_pixn_kernel<1><<<(int)grid, kPnThreads, kPnSmemBytes, st>>>(p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kPnThreads); cfg.dynamicSmemBytes = kPnSmemBytes; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, tc_conv3x3_pixn_kernel<2>, p);
        if (e != cudaSuccess) { set_error("%s: cluster launch: %s", what, cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
    }
    int rc = check_launch(what);
    return rc < 0 ? rc : (int)grid;
}

// w: P == 1: packed filter [O][9*C];  P == 2: pair-packed filter [2*O][12*C] (unetca_pack_conv3x3_pair)
static int launch_pixn(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W, int C, int O,
                       int P, float* stat_parts, cudaStream_t st, const float* ep_scale = nullptr,
                       const float* ep_shift = nullptr, float* sq_parts = nullptr) {
    PixNParams p;
    memset(&p, 0, sizeof(p));
    const int HP = H / P;
    int TW = 0, TI = 0;
    pick_tile(HP, W, 256, &TW, &TI);
    int rc;
    if ((rc = make_map5(&p.mapX, x, C, W, H, B, ldx, P, TW, TI, true)) < 0) return rc;
    const int ntaps = P == 1 ? 9 : 12;
    const long rows = P == 1 ? O : 2L * O;
    const int CL = g_pixn_cluster;
    if ((rc = make_map(&p.mapW, w, (long)ntaps * C, rows, 1, 1, ldk, rows * ldk, rows * ldk, 128 / CL, 1)) < 0) return rc;
    if ((rc = make_map5(&p.mapOut, y, O, W, H, B, ldy, P, TW, TI, false)) < 0) return rc;
    p.tilesW = ceil_div(W, TW); p.tilesI = ceil_div(HP, TI); p.nimg = B;
    p.TW = TW; p.TI = TI; p.HP = HP; p.W = W; p.P = P;
    p.twShift = 0; while ((1 << p.twShift) < TW) ++p.twShift;
    p.ntaps = ntaps; p.cchunks = C / 64;
    p.num_m_blocks = P == 1 ? O / 128 : O / 64;
    for (int t = 0; t < ntaps; ++t) {
        const int d = t / 3 - 1;                  // input row offset: -1..1 (P=1) or -1..2 relative to the even row (P=2)
        p.dw[t] = (signed char)(t % 3 - 1);
        if (P == 1) { p.par[t] = 0; p.off[t] = (signed char)d; }
        else { p.par[t] = (signed char)(d & 1); p.off[t] = (signed char)(d < 0 ? -1 : d / 2); }
    }
    p.stat_parts = stat_parts; p.N = O;
    p.ep_scale = ep_scale; p.ep_shift = ep_shift; p.sq_parts = sq_parts;
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_pixn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPnSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc_conv3x3_pixn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPnSmemBytes);
        if (e != cudaSuccess) { set_error("tc_conv3x3 (pixn): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    return launch_pixn_kernel(p, CL, st, "tc_conv3x3_fwd (pixn)");
}

// First conv in the row-pair layout (see elementwise.cu: im2col_pairs_kernel): one GEMM over the pixel-pair rows
// colp [B*(H/2)*W][64] with the pair-packed filter wp [2*O][64] through the pixels-on-N kernel (one "tap", one chunk).
static int launch_first_pairs(const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O,
                              float* stat_parts, cudaStream_t st, const float* ep_scale = nullptr,
                              const float* ep_shift = nullptr) {
    PixNParams p;
    memset(&p, 0, sizeof(p));
    const int HP = H / 2;
    int TW = 0, TI = 0;
    pick_tile(HP, W, 256, &TW, &TI);
    int rc;
    if ((rc = make_map5(&p.mapX, colp, 64, W, HP, B, 64, 1, TW, TI, true)) < 0) return rc;       // pair rows as an "image"
    const int CL = g_pixn_cluster;
    if ((rc = make_map(&p.mapW, wp, 64, 2L * O, 1, 1, 64, 2L * O * 64, 2L * O * 64, 128 / CL, 1)) < 0) return rc;
    if ((rc = make_map5(&p.mapOut, y, O, W, H, B, ldy, 2, TW, TI, false)) < 0) return rc;        // rows 2i+j of the output
    p.tilesW = ceil_div(W, TW); p.tilesI = ceil_div(HP, TI); p.nimg = B;
    p.TW = TW; p.TI = TI; p.HP = HP; p.W = W; p.P = 2;
    p.twShift = 0; while ((1 << p.twShift) < TW) ++p.twShift;
    p.ntaps = 1; p.cchunks = 1;
    p.num_m_blocks = O / 64;
    p.stat_parts = stat_parts; p.N = O;
    p.ep_scale = ep_scale; p.ep_shift = ep_shift;
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_pixn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPnSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc_conv3x3_pixn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPnSmemBytes);
        if (e != cudaSuccess) { set_error("first_pairs: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    return launch_pixn_kernel(p, CL, st, "tc_first_pairs_fwd");
}

// conv3x3 forward / dgrad, 64 -> 64 channels, H even: resident-filter row-pair kernel; w = packed filter [64][ldk >= 576]
static int launch_rp64(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W,
                       float* stat_parts, cudaStream_t st, const float* ep_scale = nullptr, const float* ep_shift = nullptr,
                       float* sq_parts = nullptr, const BwdStats* bwd = nullptr) {
    Rp64Params p;
    memset(&p, 0, sizeof(p));
    const int HP = H / 2;
    int rc;
    if ((rc = make_map5(&p.mapX, x, 64, W, H, B, ldx, 2, kRpPitch, kRpRows, true)) < 0) return rc;
    if ((rc = make_map(&p.mapW, w, 9L * 64, 64, 1, 1, ldk, 64L * ldk, 64L * ldk, 64, 1)) < 0) return rc;
    p.out = (bf16*)y; p.ldo = ldy;
    p.tilesW = ceil_div(W, kRpTW); p.tilesI = ceil_div(HP, kRpTI); p.nimg = B; p.HP = HP; p.W = W;
    p.stat_parts = stat_parts;
    p.ep_scale = ep_scale; p.ep_shift = ep_shift; p.sq_parts = sq_parts;
    if (bwd) {
        p.bwd_y = (const bf16*)bwd->y; p.bwd_ldy = bwd->ldy;
        p.bwd_scale = bwd->scale; p.bwd_shift = bwd->shift; p.bwd_mean = bwd->mean;
    }
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_conv3x3_rp64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRpSmemBytes);
        if (e != cudaSuccess) { set_error("tc_conv3x3 (rp64): cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    const long num_work = (long)p.tilesW * p.tilesI * B;
    long grid = num_work < num_sms() ? num_work : num_sms();
    if (grid < 1) grid = 1;
    tc_conv3x3_rp64_kernel<<<(int)grid, kRpThreads, kRpSmemBytes, st>>>(p);
    rc = check_launch("tc_conv3x3_fwd (rp64)");
    return rc < 0 ? rc : (int)grid;
}

static void set_taps3x3(TcParams& p) {
    for (int t = 0; t < 9; ++t) { p.dh[t] = (signed char)(t / 3 - 1); p.dw[t] = (signed char)(t % 3 - 1); p.amap[t] = 0; }
}

}  // namespace unetca

using namespace unetca;

extern "C" {

void unetca_tc_force_block_n(int n) { g_force_block_n = n; }
void unetca_tc_force_wgrad_narrow(int on) { g_wgrad_narrow = on; }
void unetca_tc_force_no_halo(int on) { g_no_halo = on; }
void unetca_tc_set_convT_wide(int on) { g_convT_wide = on; }
void unetca_tc_force_no_pixn(int on) { g_no_pixn = on; }
void unetca_tc_set_pixn_cluster(int n) { g_pixn_cluster = n == 1 ? 1 : 2; }

// conv3x3 forward / dgrad through the row-pair layout (O a multiple of 64, H even); w_pair from unetca_tc_pack_pair
int unetca_tc_conv3x3_fwd_paired(const void* x, int ldx, const void* w_pair, void* y, int ldy, int B, int H, int W, int C,
                                 int O, float* stat_parts, void* stream) {
    UNETCA_REQUIRE(C % 64 == 0 && O % 64 == 0 && O <= 1024 && H % 2 == 0,
                   "tc_conv3x3_paired: C=%d O=%d must be multiples of 64 and H=%d even", C, O, H);
    return launch_pixn(x, ldx, w_pair, 12 * C, y, ldy, B, H, W, C, O, 2, stat_parts, (cudaStream_t)stream);
}

// conv3x3 forward / dgrad for 64 -> 64 channels (H even) through the resident-filter row-pair kernel; w = packed filter [64][ldk]
int unetca_tc_conv3x3_fwd_rp64(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W,
                               float* stat_parts, void* stream) {
    UNETCA_REQUIRE(H % 2 == 0 && ldk >= 576, "tc_conv3x3_rp64: H=%d must be even, ldk=%d >= 576", H, ldk);
    return launch_rp64(x, ldx, w, ldk, y, ldy, B, H, W, stat_parts, (cudaStream_t)stream);
}

// dgrad conv3x3 (x = dY, w = dgrad-packed filter [O][9*C], y = dA) that also leaves the ReLU + BatchNorm backward
// statistics of dA in stat_parts [ret][2][O]: sum dz, sum dz*(y1 - mean), dz = dA * (scale*y1 + shift > 0).
// O % 128 == 0 (haloed kernel) or C == O == 64 with an even H (row-pair kernel); otherwise UNETCA_ERR_UNSUPPORTED.
int unetca_tc_conv3x3_dgrad_bnstats(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W,
                                    int C, int O, const void* y1, int ldy1, const float* scale, const float* shift,
                                    const float* mean, float* stat_parts, void* stream) {
    UNETCA_REQUIRE(C % 64 == 0 && O % 64 == 0 && O <= 1024 && y1 && scale && shift && mean && stat_parts,
                   "tc_conv3x3_dgrad_bnstats: C=%d O=%d", C, O);
    BwdStats bs{y1, ldy1, scale, shift, mean};
    if (O % 128 == 0) return launch_hpix(x, ldx, w, ldk, y, ldy, B, H, W, C, O, stat_parts, (cudaStream_t)stream, nullptr, nullptr, nullptr, &bs);
    if (C == 64 && O == 64 && H % 2 == 0)
        return launch_rp64(x, ldx, w, ldk, y, ldy, B, H, W, stat_parts, (cudaStream_t)stream, nullptr, nullptr, nullptr, &bs);
    set_error("tc_conv3x3_dgrad_bnstats: no fused kernel for C=%d O=%d H=%d", C, O, H);
    return UNETCA_ERR_UNSUPPORTED;
}

// conv3x3 forward / dgrad, O % 128 == 0, output channels [0, split) -> y (pixel stride ldy), [split, O) -> y2 (pixel stride ldy2)
int unetca_tc_conv3x3_fwd_split(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, void* y2, int ldy2, int split,
                                int B, int H, int W, int C, int O, float* stat_parts, void* stream) {
    UNETCA_REQUIRE(C % 64 == 0 && O % 128 == 0 && O <= 1024 && split > 0 && split < O && split % 64 == 0 && y2,
                   "tc_conv3x3_fwd_split: C=%d O=%d split=%d", C, O, split);
    return launch_hpix(x, ldx, w, ldk, y, ldy, B, H, W, C, O, stat_parts, (cudaStream_t)stream, nullptr, nullptr, nullptr, nullptr,
                       y2, ldy2, split);
}

// conv3x3 forward whose input is a channel concat kept as two dense tensors (x: channels [0, C1), x2: [C1, C)).
// O % 128 == 0: haloed kernel with w = packed filter [O][9*C]; O == 64 && C == 128: kw-stacked kernel with w = kw-stacked filter.
// scale / shift non-null: eval-mode BatchNorm + ReLU in the epilogue; sq_parts: SE squeeze sums (haloed kernel only).
int unetca_tc_conv3x3_fwd_cat(const void* x, int ldx, const void* x2, int ldx2, int C1, const void* w, void* y, int ldy, int B,
                              int H, int W, int C, int O, float* stat_parts, const float* scale, const float* shift,
                              float* sq_parts, void* stream) {
    UNETCA_REQUIRE(x2 && C % 64 == 0 && C1 > 0 && C1 < C && C1 % 64 == 0 && O % 64 == 0 && O <= 1024, "tc_conv3x3_fwd_cat: C1=%d C=%d O=%d", C1, C, O);
    XSrc2 xs{x2, ldx2, C1};
    if (O % 128 == 0)
        return launch_hpix(x, ldx, w, 9 * C, y, ldy, B, H, W, C, O, stat_parts, (cudaStream_t)stream, scale, shift, sq_parts, nullptr,
                           nullptr, 0, 0, &xs);
    if (O == 64 && C == 128 && !sq_parts)
        return launch_kw(x, ldx, w, y, ldy, B, H, W, C, stat_parts, (cudaStream_t)stream, scale, shift, &xs);
    set_error("tc_conv3x3_fwd_cat: no kernel for C=%d O=%d", C, O);
    return UNETCA_ERR_UNSUPPORTED;
}

// conv3x3 forward / dgrad for 64 output channels (C = 64 or 128) through the kw-stacked kernel
int unetca_tc_conv3x3_fwd_kw(const void* x, int ldx, const void* w_kw, void* y, int ldy, int B, int H, int W, int C,
                             float* stat_parts, void* stream) {
    UNETCA_REQUIRE(C == 64 || C == 128, "tc_conv3x3_kw: C=%d must be 64 or 128 (the filter stays resident in shared memory)", C);
    return launch_kw(x, ldx, w_kw, y, ldy, B, H, W, C, stat_parts, (cudaStream_t)stream);
}

// w_kw [9*(C/64)*64][64] from the K-major packed filter w [64][ld] (k = tap*C + c)
int unetca_tc_pack_kw(const void* w, int ld, void* w_kw, int C, void* stream) {
    UNETCA_REQUIRE(C % 64 == 0 && ld >= 9 * C, "tc_pack_kw: C=%d ld=%d", C, ld);
    const long total = 9L * (C / 64) * 64 * 64;
    pack_kw_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)w, ld, (bf16*)w_kw, C);
    return check_launch("tc_pack_kw");
}
void unetca_tc_force_no_kw(int on) { g_no_kw = on; }

// Inference forms: conv3x3 + eval-mode BatchNorm (scale, shift = running statistics folded, unetca_bn_fold_eval) + ReLU
// in the epilogue of the tcgen05 kernels, so the pre-activation tensor is never written.  layout: 0 = packed filter
// [O][9*C] (needs O % 128 == 0), 1 = pair-packed filter (O % 64 == 0, H even), 2 = kw-stacked filter (O = 64, C = 128).
// sq_parts (optional, layouts 0 and 1): [B][ret][O] per-image channel sums of the stored activations — the SE squeeze
// (UCA:65) taken in the epilogue; the caller zero-fills B * unetca_num_sms() * O floats first.  Returns the number of
// partial rows per image (> 0) or < 0 on error.
int unetca_tc_conv3x3_bnrelu_fwd(const void* x, int ldx, const void* w, int layout, void* y, int ldy, int B, int H, int W,
                                 int C, int O, const float* scale, const float* shift, float* sq_parts, void* stream) {
    UNETCA_REQUIRE(scale && shift && C % 64 == 0 && O % 64 == 0 && O <= 1024, "tc_conv3x3_bnrelu: C=%d O=%d", C, O);
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (layout == 0) {
        UNETCA_REQUIRE(O % 128 == 0, "tc_conv3x3_bnrelu: layout 0 needs O %% 128 == 0 (O=%d)", O);
        rc = launch_hpix(x, ldx, w, 9 * C, y, ldy, B, H, W, C, O, nullptr, st, scale, shift, sq_parts);
    } else if (layout == 1) {
        UNETCA_REQUIRE(H % 2 == 0, "tc_conv3x3_bnrelu: layout 1 needs an even H (H=%d)", H);
        rc = launch_pixn(x, ldx, w, 12 * C, y, ldy, B, H, W, C, O, 2, nullptr, st, scale, shift, sq_parts);
    } else if (layout == 3) {
        UNETCA_REQUIRE(H % 2 == 0 && C == 64 && O == 64, "tc_conv3x3_bnrelu: layout 3 needs C = O = 64 and an even H (C=%d O=%d H=%d)", C, O, H);
        rc = launch_rp64(x, ldx, w, 9 * C, y, ldy, B, H, W, nullptr, st, scale, shift, sq_parts);
    } else {
        UNETCA_REQUIRE(O == 64 && C == 128 && !sq_parts, "tc_conv3x3_bnrelu: layout 2 needs O = 64, C = 128, no squeeze sums (O=%d C=%d)", O, C);
        rc = launch_kw(x, ldx, w, y, ldy, B, H, W, C, nullptr, st, scale, shift);
    }
    return rc;
}
int unetca_tc_first_pairs_bnrelu_fwd(const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O,
                                     const float* scale, const float* shift, void* stream) {
    UNETCA_REQUIRE(scale && shift && O % 64 == 0 && O <= 1024 && H % 2 == 0, "tc_first_pairs_bnrelu: O=%d H=%d", O, H);
    int rc = launch_first_pairs(colp, wp, y, ldy, B, H, W, O, nullptr, (cudaStream_t)stream, scale, shift);
    return rc < 0 ? rc : 0;
}

// w_pair [2*rows][12*C] from the K-major packed filter w [rows][ld] (k = tap*C + c)
int unetca_tc_pack_pair(const void* w, int ld, void* w_pair, int rows, int C, void* stream) {
    UNETCA_REQUIRE(rows % 64 == 0 && C > 0 && ld >= 9 * C, "tc_pack_pair: rows=%d C=%d ld=%d", rows, C, ld);
    const long total = 2L * rows * 12 * C;
    pack_pair_kernel<<<ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)w, ld, (bf16*)w_pair, rows, C);
    return check_launch("tc_pack_pair");
}

// y[p][n] = sum_{tap,c} x[p+s(tap)][c] * w[n][tap*C+c]   (bf16 NHWC in/out, fp32 accumulate).
// stat_parts != null: per-CTA partial per-channel sum / sum-of-squares of the *stored* bf16 outputs,
// layout [ret][2][O]; returns the number of partial rows (>0) or <0 on error.
int unetca_tc_conv3x3_fwd(const void* x, int ldx, const void* w, int ldk, void* y, int ldy, int B, int H, int W, int C,
                          int O, float* stat_parts, void* stream) {
    UNETCA_REQUIRE(C % 64 == 0 && O % 64 == 0 && O <= 1024, "tc_conv3x3: C=%d O=%d must be multiples of 64 (O<=1024)", C, O);
    if (O % 128 == 0 && !g_no_halo && !g_force_block_n)
        return launch_hpix(x, ldx, w, ldk, y, ldy, B, H, W, C, O, stat_parts, (cudaStream_t)stream);
    if (O % 128 == 0 && O % 256 != 0 && !g_no_pixn && !g_force_block_n)
        return launch_pixn(x, ldx, w, ldk, y, ldy, B, H, W, C, O, 1, stat_parts, (cudaStream_t)stream);
    TcParams p;
    memset(&p, 0, sizeof(p));
    int TW = 0, TH = 0;
    pick_tile(H, W, 128, &TW, &TH);
    int rc;
    if ((rc = make_map(&p.mapA[0], x, C, W, H, B, ldx, (long)W * ldx, (long)H * W * ldx, TW, TH)) < 0) return rc;
    const int BN = pick_block_n(O);
    // weights as a (K, N, 1, 1) tensor, box (64, BN)
    if ((rc = make_map(&p.mapB[0], w, 9L * C, O, 1, 1, ldk, (long)O * ldk, (long)O * ldk, BN, 1)) < 0) return rc;
    if ((rc = make_map(&p.mapOut[0], y, O, W, H, B, ldy, (long)W * ldy, (long)H * W * ldy, TW, TH)) < 0) return rc;
    p.tilesW = ceil_div(W, TW); p.tilesH = ceil_div(H, TH); p.nimg = B;
    p.TW = TW; p.TH = TH; p.H = H; p.W = W;
    p.num_m_blocks = p.tilesW * p.tilesH * B;
    p.num_n_blocks = O / BN;
    p.ntaps = 9; p.cchunks = C / 64;
    set_taps3x3(p);
    p.out_chunks_per_map = O / 64;
    p.stat_parts = stat_parts; p.N = O;
    return launch_tc_n<false>(BN, p, (long)p.num_m_blocks * p.num_n_blocks, (cudaStream_t)stream, "tc_conv3x3_fwd");
}

// out[m][n] = sum_k A[m][k] * Bw[n][k], K a multiple of 64 (first conv on im2col rows)
int unetca_tc_gemm_nt(const void* A, int lda, const void* Bw, int ldb, void* out, int ldo, long M, int N, int K,
                      float* stat_parts, void* stream) {
    UNETCA_REQUIRE(K % 64 == 0 && N % 64 == 0 && N <= 1024, "tc_gemm_nt: K=%d N=%d must be multiples of 64", K, N);
    TcParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    // rows as a (K, 128-row strips, M/128 ...) map: use W = M, H = 1
    UNETCA_REQUIRE(M < (1L << 31), "tc_gemm_nt: M too large");
    if ((rc = make_map(&p.mapA[0], A, K, M, 1, 1, lda, (long)M * lda, (long)M * lda, 128, 1)) < 0) return rc;
    const int BN = pick_block_n(N);
    if ((rc = make_map(&p.mapB[0], Bw, K, N, 1, 1, ldb, (long)N * ldb, (long)N * ldb, BN, 1)) < 0) return rc;
    if ((rc = make_map(&p.mapOut[0], out, N, M, 1, 1, ldo, (long)M * ldo, (long)M * ldo, 128, 1)) < 0) return rc;
    p.tilesW = ceil_div(M, 128); p.tilesH = 1; p.nimg = 1;
    p.TW = 128; p.TH = 1; p.H = 1; p.W = (int)M;
    p.num_m_blocks = p.tilesW;
    p.num_n_blocks = N / BN;
    p.ntaps = 1; p.cchunks = K / 64;
    p.out_chunks_per_map = N / 64;
    p.stat_parts = stat_parts; p.N = N;
    return launch_tc_n<false>(BN, p, (long)p.num_m_blocks * p.num_n_blocks, (cudaStream_t)stream, "tc_gemm_nt");
}

static int g_first_wgrad_swap = 1;  // first-conv weight gradient: (j, o) on M, patch columns on N (0: patch columns on M, duplicate box)
void unetca_tc_set_first_wgrad_swap(int on) { g_first_wgrad_swap = on; }
static int g_convT_wgrad256 = 1;   // ConvTranspose weight gradient: 256 x 256 tiles where Cin % 256 == 0 (0: generic 128 x 256 kernel)
void unetca_tc_set_convT_wgrad256(int on) { g_convT_wgrad256 = on; }
static int g_convT_pix = 1;        // ConvTranspose forward through the dedicated pixels-on-N kernel (0: generic pixels-on-M kernel)
void unetca_tc_set_convT_pix(int on) { g_convT_pix = on; }

static int launch_convT_fwd_pix(const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B, int h, int wd,
                                int Cin, int Cout, cudaStream_t st) {
    CtfParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    if ((rc = make_map(&p.mapX, x, Cin, wd, h, B, ldx, (long)wd * ldx, (long)h * wd * ldx, kHpTW, kHpTH)) < 0) return rc;
    if ((rc = make_map(&p.mapW, w, Cin, 4L * Cout, 1, 1, Cin, 4L * Cout * Cin, 4L * Cout * Cin, 128, 1)) < 0) return rc;
    p.out = (bf16*)out; p.ldo = ldo; p.bias = bias;
    p.tilesW = ceil_div(wd, kHpTW); p.tilesH = ceil_div(h, kHpTH); p.nimg = B; p.h = h; p.w = wd;
    p.cchunks = Cin / 64; p.num_m_blocks = 4 * Cout / 128; p.Cout = Cout;
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_convT_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCtSmemBytes);
        if (e != cudaSuccess) { set_error("tc_convT_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    const long num_work = (long)p.tilesW * p.tilesH * B * p.num_m_blocks;
    long grid = num_work < num_sms() ? num_work : num_sms();
    if (grid < 1) grid = 1;
    tc_convT_fwd_kernel<<<(int)grid, kCtThreads, kCtSmemBytes, st>>>(p);
    return check_launch("tc_convT_fwd (pixels on N)");
}

// ConvTranspose2d k2 s2: out[b,2i+d,2j+e,o] = sum_c x[b,i,j,c] * w[(d*2+e)*Cout+o][c] + bias[o]
int unetca_tc_convT_fwd(const void* x, int ldx, const void* w, const float* bias, void* out, int ldo, int B, int h,
                        int wd, int Cin, int Cout, void* stream) {
    UNETCA_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_convT: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
    // tiles of 8 x 32 input pixels: below half a tile of rows the generic kernel's 128-pixel tiles waste less
    if (g_convT_pix && bias && h >= 16 && wd >= 8) {
        int r = launch_convT_fwd_pix(x, ldx, w, bias, out, ldo, B, h, wd, Cin, Cout, (cudaStream_t)stream);
        return r < 0 ? r : 0;
    }
    TcParams p;
    memset(&p, 0, sizeof(p));
    int TW = 0, TH = 0;
    pick_tile(h, wd, 128, &TW, &TH);
    int rc;
    if ((rc = make_map(&p.mapA[0], x, Cin, wd, h, B, ldx, (long)wd * ldx, (long)h * wd * ldx, TW, TH)) < 0) return rc;
    // one N block may span several (d,e) sub-pixel maps: 4*Cout is always a multiple of 256, the width at which an
    // SS-mode MMA reaches full rate, and the activation tile is then read once instead of once per sub-pixel
    const int BN = g_convT_wide ? pick_block_n(4 * Cout) : pick_block_n(Cout);
    if ((rc = make_map(&p.mapB[0], w, Cin, 4L * Cout, 1, 1, Cin, 4L * Cout * Cin, 4L * Cout * Cin, BN, 1)) < 0) return rc;
    for (int de = 0; de < 4; ++de) {
        const bf16* base = (const bf16*)out + ((long)(de >> 1) * 2 * wd + (de & 1)) * ldo;
        if ((rc = make_map(&p.mapOut[de], base, Cout, wd, h, B, 2L * ldo, 4L * wd * ldo, 4L * h * wd * ldo, TW, TH)) < 0) return rc;
    }
    p.tilesW = ceil_div(wd, TW); p.tilesH = ceil_div(h, TH); p.nimg = B;
    p.TW = TW; p.TH = TH; p.H = h; p.W = wd;
    p.num_m_blocks = p.tilesW * p.tilesH * B;
    p.num_n_blocks = 4 * Cout / BN;
    p.ntaps = 1; p.cchunks = Cin / 64;
    p.out_chunks_per_map = Cout / 64;
    p.bias = bias; p.bias_mod = Cout;
    p.N = 4 * Cout;
    int r2 = launch_tc_n<false>(BN, p, (long)p.num_m_blocks * p.num_n_blocks, (cudaStream_t)stream, "tc_convT_fwd");
    return r2 < 0 ? r2 : 0;
}

// dx[b,i,j,c] = sum_{de,o} dout[b,2i+d,2j+e,o] * wdg[c][de*Cout+o]
int unetca_tc_convT_dgrad(const void* dout, int ldd, const void* wdg, void* dx, int ldx, int B, int h, int wd, int Cin,
                          int Cout, void* stream) {
    UNETCA_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_convT: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
    TcParams p;
    memset(&p, 0, sizeof(p));
    int TW = 0, TH = 0;
    pick_tile(h, wd, 128, &TW, &TH);
    int rc;
    for (int de = 0; de < 4; ++de) {
        const bf16* base = (const bf16*)dout + ((long)(de >> 1) * 2 * wd + (de & 1)) * ldd;
        if ((rc = make_map(&p.mapA[de], base, Cout, wd, h, B, 2L * ldd, 4L * wd * ldd, 4L * h * wd * ldd, TW, TH)) < 0) return rc;
        p.amap[de] = (signed char)de;
    }
    const int BN = pick_block_n(Cin);
    if ((rc = make_map(&p.mapB[0], wdg, 4L * Cout, Cin, 1, 1, 4L * Cout, 4L * Cout * Cin, 4L * Cout * Cin, BN, 1)) < 0) return rc;
    if ((rc = make_map(&p.mapOut[0], dx, Cin, wd, h, B, ldx, (long)wd * ldx, (long)h * wd * ldx, TW, TH)) < 0) return rc;
    p.tilesW = ceil_div(wd, TW); p.tilesH = ceil_div(h, TH); p.nimg = B;
    p.TW = TW; p.TH = TH; p.H = h; p.W = wd;
    p.num_m_blocks = p.tilesW * p.tilesH * B;
    p.num_n_blocks = Cin / BN;
    p.ntaps = 4; p.cchunks = Cout / 64;
    p.out_chunks_per_map = Cin / 64;
    p.N = Cin;
    int r2 = launch_tc_n<false>(BN, p, (long)p.num_m_blocks * p.num_n_blocks, (cudaStream_t)stream, "tc_convT_dgrad");
    return r2 < 0 ? r2 : 0;
}

// Split-K factor for a persistent weight-gradient kernel: `items` equal work items per split walk the grid in rounds of
// num_sms() CTAs, so the kernel takes ceil(items * s / num_sms()) rounds of (1 / s) of the K range each.  Around the
// target number of waves, pick the s with the least total time = rounds / s — i.e. the one whose last round is full
// (items = 16, 148 SMs: s = 19 -> 3 rounds of 1/19 = 0.158; s = 18 -> 2 rounds of 1/18 = 0.111 of the serial time).
static long fit_split(long items, int waves, long smax) {
    const long G = num_sms();
    long target = ((long)waves * G + items - 1) / items;
    if (target > smax) target = smax;
    if (target < 1) target = 1;
    long lo = target - target / 3, hi = target + target / 4;
    if (lo < 1) lo = 1;
    if (hi > smax) hi = smax;
    long best = target;
    double best_t = 1e30;
    for (long s = lo; s <= hi; ++s) {
        const long rounds = (items * s + G - 1) / G;
        const double t = (double)rounds / (double)s;
        if (t < best_t * (1.0 - 1e-9)) { best_t = t; best = s; }
    }
    return best;
}

static int pick_nsplit(long tiles, int ktiles, long stride_floats, long ws_floats) {
    long smax = ktiles / 4;
    if (ws_floats / stride_floats < smax) smax = ws_floats / stride_floats;
    if (smax < 1) return (int)(ws_floats / stride_floats);        // 0 when the workspace cannot hold one slice
    return (int)fit_split(tiles, 2, smax);
}

// ws[z][o][tap*C+c] = sum_{p in split z} dy[p][o] * x[p+s(tap)][c]; returns nsplit.  Halo-reuse kernel.
static int tc_conv3x3_wgrad_impl(const void* dy, int lddy, const void* x, int ldx, const XSrc2* xs2, float* ws, long ws_floats,
                                 int B, int H, int W, int C, int O, void* stream) {
    UNETCA_REQUIRE(C % 64 == 0 && O % 64 == 0, "tc_conv3x3_wgrad: C=%d O=%d must be multiples of 64", C, O);
    Wg3Params p;
    memset(&p, 0, sizeof(p));
    int rc;
    const bool wide = (C % 128 == 0) && (O % 128 == 0) && g_wgrad_narrow != 1;
    const bool wide256 = wide && (O % 256 == 0) && g_wgrad_narrow != 2;
    const bool rowpair = !wide && g_wgrad_narrow == 0;
    const int xrows = wide256 ? kWg3XRows : wide ? kWgTH : rowpair ? kWg4XRows : kWgTH + 2;
    const int C1 = xs2 ? xs2->xsplit : C;
    if ((rc = make_map(&p.mapX, x, C1, W, H, B, ldx, (long)W * ldx, (long)H * W * ldx, kWgPitch, xrows)) < 0) return rc;
    if (xs2) {
        if ((rc = make_map(&p.mapX2, xs2->x2, C - C1, W, H, B, xs2->ldx2, (long)W * xs2->ldx2, (long)H * W * xs2->ldx2, kWgPitch,
                           xrows)) < 0) return rc;
        p.xsplit = C1;
    }
    if ((rc = make_map(&p.mapDY, dy, O, W, H, B, lddy, (long)W * lddy, (long)H * W * lddy, kWgTW, kWgTH)) < 0) return rc;
    p.cchunks = C / 64; p.oblocks = O / 64;
    p.tilesW = ceil_div(W, kWgTW); p.tilesH = ceil_div(H, kWgTH); p.nimg = B;
    p.ktiles_total = p.tilesW * p.tilesH * B;
    p.split_stride = (long long)O * 9 * C;
    const long items = wide256 ? (long)(C / 128) * (O / 256) * 5 : wide ? (long)(C / 128) * (O / 128) * 3
                       : rowpair ? (long)p.cchunks * p.oblocks * 2 : (long)p.cchunks * p.oblocks;
    const int slices_per_split = rowpair ? 2 : 1;        // the row-pair kernel writes the j = 0 / j = 1 halves separately
    long smax = p.ktiles_total / 16;
    if (smax < 1) smax = 1;
    if (ws_floats / (p.split_stride * slices_per_split) < smax) smax = ws_floats / (p.split_stride * slices_per_split);
    if (smax < 1) { set_error("tc_conv3x3_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
    const long ns = g_wgrad_waves > 0 ? fit_split(items, g_wgrad_waves, smax) : 1;
    p.nsplit = (int)ns;
    p.ws = ws; p.ldn = 9 * C; p.C = C;
    static bool attr_done_dev[kMaxDevices] = {};
    bool& attr_done = attr_done_dev[device_slot()];          // cudaFuncSetAttribute is per device
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(tc_wgrad3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc_wgrad3x3_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWg2SmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc_wgrad3x3_wide256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWg3SmemBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(tc_wgrad3x3_rowpair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWg4SmemBytes);
        if (e != cudaSuccess) { set_error("tc_conv3x3_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
        attr_done = true;
    }
    long grid = items * ns < num_sms() ? items * ns : num_sms();
    if (wide256) tc_wgrad3x3_wide256_kernel<<<(int)grid, kTcThreads, kWg3SmemBytes, (cudaStream_t)stream>>>(p);
    else if (wide) tc_wgrad3x3_wide_kernel<<<(int)grid, kTcThreads, kWg2SmemBytes, (cudaStream_t)stream>>>(p);
    else if (rowpair) tc_wgrad3x3_rowpair_kernel<<<(int)grid, kTcThreads, kWg4SmemBytes, (cudaStream_t)stream>>>(p);
    else tc_wgrad3x3_kernel<<<(int)grid, kTcThreads, kWgSmemBytes, (cudaStream_t)stream>>>(p);
    rc = check_launch("tc_conv3x3_wgrad");
    return rc < 0 ? rc : p.nsplit * slices_per_split;
}

int unetca_tc_conv3x3_wgrad(const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats, int B, int H,
                            int W, int C, int O, void* stream) {
    return tc_conv3x3_wgrad_impl(dy, lddy, x, ldx, nullptr, ws, ws_floats, B, H, W, C, O, stream);
}
// the same with the activation given as two dense tensors: channels [0, C1) in x, [C1, C) in x2 (C1 a multiple of 64)
int unetca_tc_conv3x3_wgrad_cat(const void* dy, int lddy, const void* x, int ldx, const void* x2, int ldx2, int C1, float* ws,
                                long ws_floats, int B, int H, int W, int C, int O, void* stream) {
    UNETCA_REQUIRE(x2 && C1 > 0 && C1 < C && C1 % 64 == 0, "tc_conv3x3_wgrad_cat: C1=%d C=%d", C1, C);
    XSrc2 xs{x2, ldx2, C1};
    return tc_conv3x3_wgrad_impl(dy, lddy, x, ldx, &xs, ws, ws_floats, B, H, W, C, O, stream);
}

// the generic (one box per tap) kernel, kept for cross-checks
int unetca_tc_conv3x3_wgrad_generic(const void* dy, int lddy, const void* x, int ldx, float* ws, long ws_floats, int B, int H,
                            int W, int C, int O, void* stream) {
    UNETCA_REQUIRE(C % 64 == 0 && O % 64 == 0, "tc_conv3x3_wgrad: C=%d O=%d must be multiples of 64", C, O);
    TcParams p;
    memset(&p, 0, sizeof(p));
    int TW = 0, TH = 0;
    pick_tile(H, W, 64, &TW, &TH);
    int rc;
    if ((rc = make_map(&p.mapA[0], x, C, W, H, B, ldx, (long)W * ldx, (long)H * W * ldx, TW, TH)) < 0) return rc;
    if ((rc = make_map(&p.mapB[0], dy, O, W, H, B, lddy, (long)W * lddy, (long)H * W * lddy, TW, TH)) < 0) return rc;
    const int BN = pick_block_n(O);
    p.tilesW = ceil_div(W, TW); p.tilesH = ceil_div(H, TH); p.nimg = B;
    p.TW = TW; p.TH = TH; p.H = H; p.W = W;
    set_taps3x3(p);
    p.a_chunks = 9 * (C / 64); p.a_cchunks = C / 64; p.b_chunks_per_map = O / 64;
    p.num_m_blocks = (p.a_chunks + 1) / 2;
    p.num_n_blocks = O / BN;
    p.ktiles_total = p.tilesW * p.tilesH * B;
    p.split_stride = (long long)O * 9 * C;
    p.nsplit = pick_nsplit((long)p.num_m_blocks * p.num_n_blocks, p.ktiles_total, p.split_stride, ws_floats);
    if (p.nsplit < 1) { set_error("tc_conv3x3_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
    p.ws = ws; p.ldn = 9 * C; p.store_transposed = 1; p.m_valid = 9 * C;
    int r2 = launch_tc_n<true>(BN, p, (long)p.num_m_blocks * p.num_n_blocks * p.nsplit, (cudaStream_t)stream, "tc_conv3x3_wgrad");
    return r2 < 0 ? r2 : p.nsplit;
}

// ws[z][m][n] = sum_k A[k][m] * Bm[k][n]   (A: [K][M], Bm: [K][N], both channel-contiguous; M, N multiples of 64).
// First-conv wgrad: A = dY [npix][O], Bm = im2col rows [npix][Kpad] -> ws[z][O][Kpad].  returns nsplit
int unetca_tc_gemm_tn(const void* A, int lda, const void* Bm, int ldb, float* ws, long ws_floats, int M, int N, long K,
                      void* stream) {
    UNETCA_REQUIRE(M % 64 == 0 && N % 64 == 0, "tc_gemm_tn: M=%d N=%d must be multiples of 64", M, N);
    UNETCA_REQUIRE(K < (1L << 31), "tc_gemm_tn: K too large");
    TcParams p;
    memset(&p, 0, sizeof(p));
    int rc;
    // accumulator rows <- Bm's channels (n), accumulator columns <- A's channels (m); stored transposed
    if ((rc = make_map(&p.mapA[0], Bm, N, K, 1, 1, ldb, K * ldb, K * ldb, 64, 1)) < 0) return rc;
    if ((rc = make_map(&p.mapB[0], A, M, K, 1, 1, lda, K * lda, K * lda, 64, 1)) < 0) return rc;
    const int BN = pick_block_n(M);
    p.tilesW = ceil_div(K, 64); p.tilesH = 1; p.nimg = 1;
    p.TW = 64; p.TH = 1; p.H = 1; p.W = (int)K;
    p.a_chunks = N / 64; p.a_cchunks = N / 64; p.b_chunks_per_map = M / 64;
    p.num_m_blocks = (p.a_chunks + 1) / 2;
    p.num_n_blocks = M / BN;
    p.ktiles_total = p.tilesW;
    p.split_stride = (long long)M * N;
    p.nsplit = pick_nsplit((long)p.num_m_blocks * p.num_n_blocks, p.ktiles_total, p.split_stride, ws_floats);
    if (p.nsplit < 1) { set_error("tc_gemm_tn: workspace too small"); return UNETCA_ERR_WORKSPACE; }
    p.ws = ws; p.ldn = N; p.store_transposed = 1; p.m_valid = N;
    int r2 = launch_tc_n<true>(BN, p, (long)p.num_m_blocks * p.num_n_blocks * p.nsplit, (cudaStream_t)stream, "tc_gemm_tn");
    return r2 < 0 ? r2 : p.nsplit;
}

// first conv forward through the row-pair layout: colp from unetca_im2col_pairs, wp from unetca_pack_first_pairs
int unetca_tc_first_pairs_fwd(const void* colp, const void* wp, void* y, int ldy, int B, int H, int W, int O,
                              float* stat_parts, void* stream) {
    UNETCA_REQUIRE(O % 64 == 0 && O <= 1024 && H % 2 == 0, "tc_first_pairs_fwd: O=%d H=%d", O, H);
    return launch_first_pairs(colp, wp, y, ldy, B, H, W, O, stat_parts, (cudaStream_t)stream);
}

// ws[z][k (64 rows)][(j, o)] = sum over pixel pairs of colp[pair][k] * dy[(2i+j, x)][o]; returns nsplit
int unetca_tc_first_pairs_wgrad(const void* dy, int lddy, const void* colp, float* ws, long ws_floats, int B, int H, int W,
                                int O, void* stream) {
    UNETCA_REQUIRE(O % 64 == 0 && H % 2 == 0, "tc_first_pairs_wgrad: O=%d H=%d", O, H);
    TcParams p;
    memset(&p, 0, sizeof(p));
    const int HP = H / 2;
    int TW = 0, TH = 0;
    pick_tile(HP, W, 64, &TW, &TH);
    int rc;
    if ((rc = make_map(&p.mapA[0], colp, 64, W, HP, B, 64, (long)W * 64, (long)HP * W * 64, TW, TH)) < 0) return rc;
    for (int j = 0; j < 2; ++j) {
        const bf16* base = (const bf16*)dy + (long)j * W * lddy;
        if ((rc = make_map(&p.mapB[j], base, O, W, HP, B, lddy, 2L * W * lddy, (long)H * W * lddy, TW, TH)) < 0) return rc;
    }
    p.tilesW = ceil_div(W, TW); p.tilesH = ceil_div(HP, TH); p.nimg = B;
    p.TW = TW; p.TH = TH; p.H = HP; p.W = W;
    p.ktiles_total = p.tilesW * p.tilesH * B;
    p.split_stride = 64LL * 2 * O;
    p.ws = ws; p.ldn = 2 * O;
    if (O == 64 && g_first_wgrad_swap) {
        // (j, o) on M, the 64 patch columns on N: the two gradient boxes fill all 128 MMA rows.  With the patch columns on M
        // (below) the second half of every A tile is a duplicate box — a quarter of the shared-memory fill and half of every
        // MMA wasted on a kernel whose bound is the 3.2 GB it streams.  Output transposed on store, so ws keeps its layout.
        CUtensorMap col = p.mapA[0];
        p.mapA[0] = p.mapB[0]; p.mapA[1] = p.mapB[1]; p.mapB[0] = col;
        p.amap[0] = 0; p.amap[1] = 1;
        p.a_chunks = 2; p.a_cchunks = 1; p.b_chunks_per_map = 1;
        p.num_m_blocks = 1; p.num_n_blocks = 1;
        p.nsplit = pick_nsplit(1, p.ktiles_total, p.split_stride, ws_floats);
        if (p.nsplit < 1) { set_error("tc_first_pairs_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
        p.store_transposed = 1; p.m_valid = 128;
        int r3 = launch_tc_n<true>(64, p, (long)p.nsplit, (cudaStream_t)stream, "tc_first_pairs_wgrad");
        return r3 < 0 ? r3 : p.nsplit;
    }
    const int BN = pick_block_n(2 * O);
    p.a_chunks = 1; p.a_cchunks = 1; p.b_chunks_per_map = O / 64;
    p.num_m_blocks = 1;
    p.num_n_blocks = 2 * O / BN;
    p.nsplit = pick_nsplit((long)p.num_n_blocks, p.ktiles_total, p.split_stride, ws_floats);
    if (p.nsplit < 1) { set_error("tc_first_pairs_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
    p.store_transposed = 0; p.m_valid = 64;
    int r2 = launch_tc_n<true>(BN, p, (long)p.num_n_blocks * p.nsplit, (cudaStream_t)stream, "tc_first_pairs_wgrad");
    return r2 < 0 ? r2 : p.nsplit;
}

// ws[z][cin][de*Cout+o] = sum_p x[p][cin] * dout[(p,de)][o]; returns nsplit
int unetca_tc_convT_wgrad(const void* x, int ldx, const void* dout, int ldd, float* ws, long ws_floats, int B, int h,
                          int wd, int Cin, int Cout, void* stream) {
    UNETCA_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0, "tc_convT_wgrad: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
    TcParams p;
    memset(&p, 0, sizeof(p));
    int TW = 0, TH = 0;
    pick_tile(h, wd, 64, &TW, &TH);
    int rc;
    if ((rc = make_map(&p.mapA[0], x, Cin, wd, h, B, ldx, (long)wd * ldx, (long)h * wd * ldx, TW, TH)) < 0) return rc;
    for (int de = 0; de < 4; ++de) {
        const bf16* base = (const bf16*)dout + ((long)(de >> 1) * 2 * wd + (de & 1)) * ldd;
        if ((rc = make_map(&p.mapB[de], base, Cout, wd, h, B, 2L * ldd, 4L * wd * ldd, 4L * h * wd * ldd, TW, TH)) < 0) return rc;
    }
    const int BN = g_convT_wide ? pick_block_n(4 * Cout) : pick_block_n(Cout);     // N blocks may span the (d,e) maps
    p.tilesW = ceil_div(wd, TW); p.tilesH = ceil_div(h, TH); p.nimg = B;
    p.TW = TW; p.TH = TH; p.H = h; p.W = wd;
    p.a_chunks = Cin / 64; p.a_cchunks = Cin / 64; p.b_chunks_per_map = Cout / 64;
    if (g_convT_wgrad256 && Cin % 256 == 0 && g_convT_wide && !g_force_block_n) {
        // 256 x 256 output tiles (two accumulators): a third less shared-memory fill per MMA than the generic kernel
        p.num_m_blocks = Cin / 256;
        p.num_n_blocks = 4 * Cout / 256;
        p.ktiles_total = p.tilesW * p.tilesH * B;
        p.split_stride = (long long)Cin * 4 * Cout;
        p.nsplit = pick_nsplit((long)p.num_m_blocks * p.num_n_blocks, p.ktiles_total, p.split_stride, ws_floats);
        if (p.nsplit < 1) { set_error("tc_convT_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
        p.ws = ws; p.ldn = 4 * Cout; p.store_transposed = 0; p.m_valid = Cin;
        static bool attr_done_dev[kMaxDevices] = {};
        bool& attr_done = attr_done_dev[device_slot()];
        if (!attr_done) {
            cudaError_t e = cudaFuncSetAttribute(tc_convT_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCwSmemBytes);
            if (e != cudaSuccess) { set_error("tc_convT_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UNETCA_ERR_CUDA; }
            attr_done = true;
        }
        const long num_work = (long)p.num_m_blocks * p.num_n_blocks * p.nsplit;
        long grid = num_work < num_sms() ? num_work : num_sms();
        if (grid < 1) grid = 1;
        tc_convT_wgrad_kernel<<<(int)grid, kTcThreads, kCwSmemBytes, (cudaStream_t)stream>>>(p);
        int r3 = check_launch("tc_convT_wgrad (256 x 256)");
        return r3 < 0 ? r3 : p.nsplit;
    }
    p.num_m_blocks = (p.a_chunks + 1) / 2;
    p.num_n_blocks = 4 * Cout / BN;
    p.ktiles_total = p.tilesW * p.tilesH * B;
    p.split_stride = (long long)Cin * 4 * Cout;
    p.nsplit = pick_nsplit((long)p.num_m_blocks * p.num_n_blocks, p.ktiles_total, p.split_stride, ws_floats);
    if (p.nsplit < 1) { set_error("tc_convT_wgrad: workspace too small"); return UNETCA_ERR_WORKSPACE; }
    p.ws = ws; p.ldn = 4 * Cout; p.store_transposed = 0; p.m_valid = Cin;
    int r2 = launch_tc_n<true>(BN, p, (long)p.num_m_blocks * p.num_n_blocks * p.nsplit, (cudaStream_t)stream, "tc_convT_wgrad");
    return r2 < 0 ? r2 : p.nsplit;
}

}  // extern "C"
