// umma_shift_probe.cu — hardware probe (dev tool, not part of the library): does a tcgen05 shared-memory descriptor
// whose start address is shifted by a whole number of 128-byte rows (not a multiple of the 1024-byte swizzle atom),
// or whose stride-byte-offset is not a multiple of 1024, still address the SWIZZLE_128B data that TMA wrote?
// If yes, the 9 taps of a 3x3 conv can be issued from ONE haloed activation tile in shared memory.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I insar-unet-ca_b200/csrc tools/umma_shift_probe.cu -o tools/umma_shift_probe
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace unetca;
namespace unetca { void set_error(const char*, ...) {} int check_launch(const char*) { return 0; } int num_sms() { return 148; } }

struct Args {
    CUtensorMap mapA, mapB;
    int mode;        // 0: K-major A with row shift / pitch;  1: MN-major, B operand shifted along K
    int r0;          // row shift
    int bo;          // base_offset field value
    int sbo;         // stride byte offset for the shifted operand
    float* out;      // [128][64]
};

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t bo) {
    return make_smem_desc(saddr, lbo, sbo) | ((uint64_t)(bo & 7) << 49);
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Args a) {
    extern __shared__ uint8_t raw[];
    uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = sm;                 // 512 rows x 128 B = 64 KB
    uint8_t* sB = sm + 65536;         // 256 rows x 128 B = 32 KB
    uint64_t* bar = (uint64_t*)(sm + 65536 + 32768);
    uint32_t* slot = (uint32_t*)(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<64>(slot);
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar[0], 65536 + 32768);
        tma_load_4d(&a.mapA, &bar[0], sA, 0, 0, 0, 0);
        tma_load_4d(&a.mapA, &bar[0], sA + 32768, 0, 256, 0, 0);
        tma_load_4d(&a.mapB, &bar[0], sB, 0, 0, 0, 0);
        mbar_wait(&bar[0], 0);
        tcgen05_fence_after();
        if (a.mode == 0) {
            constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
            for (int k = 0; k < 4; ++k) {
                uint64_t da = desc_bo(smem_u32(sA) + a.r0 * 128 + k * 32, 16, a.sbo, a.bo);
                uint64_t db = make_smem_desc(smem_u32(sB) + k * 32, 16, 1024);
                umma_bf16(tmem, da, db, idesc, k != 0);
            }
        } else if (a.mode == 2) {
            // MN-major A with M = 128 made of TWO ROW-SHIFTED views of the same tile: LBO = lbo_rows * 128 bytes
            constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
            for (int k = 0; k < 4; ++k) {
                uint64_t da = make_smem_desc(smem_u32(sA) + (a.r0 + k * 16) * 128, a.sbo /* = LBO here */, 1024);
                uint64_t db = make_smem_desc(smem_u32(sB) + k * 2048, 8192, 1024);
                umma_bf16(tmem, da, db, idesc, k != 0);
            }
        } else {
            constexpr uint32_t idesc = make_idesc(128, 64, 1, 1);
            for (int k = 0; k < 4; ++k) {      // K = 64 rows
                uint64_t da = make_smem_desc(smem_u32(sA) + k * 2048, 32768, 1024);           // M = 128: 2 boxes, LBO = 32 KB
                uint64_t db = desc_bo(smem_u32(sB) + (a.r0 + k * 16) * 128, 8192, a.sbo, a.bo);
                umma_bf16(tmem, da, db, idesc, k != 0);
            }
        }
        umma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tcgen05_fence_after();
    uint32_t v[32];
    for (int h = 0; h < 2; ++h) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + h * 32, v);
        tmem_wait_ld();
        for (int i = 0; i < 32; ++i) a.out[(warp * 32 + lane) * 64 + h * 32 + i] = __uint_as_float(v[i]);
    }
    tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void mk(EncFn enc, CUtensorMap* m, void* base, int rows, int box_rows) {
    cuuint64_t dims[4] = {64, (cuuint64_t)rows, 1, 1};
    cuuint64_t str[3] = {128, (cuuint64_t)rows * 128, (cuuint64_t)rows * 128};
    cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
    void* fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    EncFn enc = (EncFn)fp;
    const int RA = 512, RB = 256;
    std::vector<float> A(RA * 64), Bm(RB * 64);
    std::vector<__nv_bfloat16> Ah(RA * 64), Bh(RB * 64);
    srand(1);
    for (int i = 0; i < RA * 64; ++i) { A[i] = (float)(rand() % 9 - 4); Ah[i] = __float2bfloat16(A[i]); }
    for (int i = 0; i < RB * 64; ++i) { Bm[i] = (float)(rand() % 7 - 3); Bh[i] = __float2bfloat16(Bm[i]); }
    __nv_bfloat16 *dA, *dB; float* dO;
    cudaMalloc(&dA, RA * 128); cudaMalloc(&dB, RB * 128); cudaMalloc(&dO, 128 * 64 * 4);
    cudaMemcpy(dA, Ah.data(), RA * 128, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bh.data(), RB * 128, cudaMemcpyHostToDevice);
    Args a;
    mk(enc, &a.mapA, dA, RA, 256);
    mk(enc, &a.mapB, dB, RB, 256);
    a.out = dO;
    const int smem = 1024 + 65536 + 32768 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> O(128 * 64);
    auto run = [&](int mode, int r0, int bo, int sbo) {
        a.mode = mode; a.r0 = r0; a.bo = bo; a.sbo = sbo;
        cudaMemset(dO, 0xff, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d r0 %d bo %d sbo %d: CUDA error %s\n", mode, r0, bo, sbo, cudaGetErrorString(e)); exit(2); }
        cudaMemcpy(O.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
                double ref = 0;
                if (mode == 0) {
                    const int row = r0 + (m / 8) * (sbo / 128) + (m % 8);
                    for (int k = 0; k < 64; ++k) ref += A[row * 64 + k] * Bm[n * 64 + k];
                } else if (mode == 2) {
                    // D[m][n] = sum_k A[r0 + (m/64)*lbo_rows + k][m%64] * B[k][n]
                    for (int k = 0; k < 64; ++k) ref += A[(r0 + (m / 64) * (sbo / 128) + k) * 64 + (m % 64)] * Bm[k * 64 + n];
                } else {
                    // D[m][n] = sum_k Aop[k][m] * Bop[r0 + k][n]; Aop box0 = rows 0..255 (m < 64 -> col m), box1 = rows 256.. (col m-64)
                    for (int k = 0; k < 64; ++k) {
                        const double av = m < 64 ? A[k * 64 + m] : A[(256 + k) * 64 + (m - 64)];
                        const int kr = r0 + (k / 8) * (sbo / 128) + (k % 8);
                        ref += av * Bm[kr * 64 + n];
                    }
                }
                if (fabs(ref - O[m * 64 + n]) > 1e-3) ++bad;
            }
        printf("mode %d (%s) r0=%2d base_offset=%d sbo|lbo=%4d : %s (%d / 8192 wrong)\n", mode,
               mode == 2 ? "MN-major A, M-blocks = row-shifted views (LBO)" : mode ? "MN-major B shifted in K" : "K-major A shifted rows",
               r0, bo, sbo, bad ? "MISMATCH" : "ok", bad);
        return bad;
    };
    for (int lbo_rows : {1, 2, 16, 18, 0})
        for (int r0 : {0, 1, 19, 38}) run(2, r0, 0, lbo_rows * 128);
    const int shifts[] = {0, 8, 1, 2, 3, 5, 9, 10, 11, 18, 23};
    for (int mode = 0; mode < 2; ++mode)
        for (int r0 : shifts) {
            run(mode, r0, 0, 1024);
            if (r0 & 7) run(mode, r0, r0 & 7, 1024);
        }
    // non-1024 stride between 8-row groups (haloed tile pitch of 10 / 12 / 18 rows), K-major
    for (int pitch : {10, 12, 16, 18})
        for (int r0 : {0, 1, 11}) {
            run(0, r0, 0, pitch * 128);
            if (r0 & 7) run(0, r0, r0 & 7, pitch * 128);
        }
    return 0;
}
