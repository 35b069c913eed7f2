// umma_rate_probe.cu — how many cycles does one tcgen05.mma (kind::f16, bf16, K=16, operands in shared memory)
// take as a function of N, for cta_group::1 (M=128) and cta_group::2 (M=256 over a CTA pair)?  No loads, no epilogue:
// one thread issues `iters` MMAs over K-major SWIZZLE_128B tiles that already sit in shared memory.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I insar-unet-ca_b200/csrc tools/umma_rate_probe.cu -o tools/umma_rate_probe
#include "tc_ptx.cuh"
#include <cstdio>
#include <vector>
using namespace unetca;

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// mode: 0 = K-major A and B, 1 = MN-major A and B
template <int CTAS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int mode, int a_stride, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // fill 192 KB with a non-trivial bf16 pattern (small values)
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += blockDim.x) w[i] = 0x3c003c00u ^ ((i * 2654435761u) & 0x007f007fu);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { if (CTAS == 1) tmem_alloc<512>(&tmem_slot); else tmem_alloc2<512>(&tmem_slot); }
    tcgen05_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const bool leader = CTAS == 1 || cluster_ctarank() == 0;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        if (leader) {
            const uint32_t idesc = make_idesc(128 * CTAS, N, mode, mode);
            const uint32_t a0 = smem_u32(smem), b0 = a0 + 96 * 1024;
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const int st = (i >> 2) % 5, k = i & 3;
                uint64_t da, db;
                if (mode == 0) {
                    da = make_smem_desc(a0 + st * 16384 + k * 32, 16, a_stride);
                    db = make_smem_desc(b0 + st * 16384 + k * 32, 16, 1024);
                } else {
                    da = make_smem_desc(a0 + st * 16384 + k * 2048, 8192, 1024);
                    db = make_smem_desc(b0 + st * 16384 + k * 2048, 8192, 1024);
                }
                if (CTAS == 1) umma_bf16(tmem_base + (i & 1) * 256, da, db, idesc, 1);
                else umma_bf16_2cta(tmem_base + (i & 1) * 256, da, db, idesc, 1);
            }
            if (CTAS == 1) umma_commit(&bar); else umma_commit_2cta(&bar);
        }
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (leader) cycles_out[blockIdx.x] = t1 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CTAS == 2) cluster_sync_all();
    tcgen05_fence_after();
    if (warp == 0) { if (CTAS == 1) tmem_dealloc<512>(tmem_base); else tmem_dealloc2<512>(tmem_base); }
}

int main() {
    int nsm = 148;
    long long* d;
    cudaMalloc(&d, 1024 * sizeof(long long));
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 20000;
    for (int ctas = 1; ctas <= 2; ++ctas)
        for (int mode = 0; mode < 2; ++mode)
            for (int stride : {1024, 1280}) {
                if (mode == 1 && stride != 1024) continue;
                for (int N : {64, 128, 192, 256}) {
                    for (int grid : {2, nsm}) {
                        cudaMemset(d, 0, 1024 * sizeof(long long));
                        cudaEvent_t e0, e1;
                        cudaEventCreate(&e0); cudaEventCreate(&e1);
                        cudaEventRecord(e0);
                        if (ctas == 1) rate_kernel<1><<<grid, 128, smem>>>(N, iters, mode, stride, d);
                        else {
                            cudaLaunchConfig_t cfg = {};
                            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                            cudaLaunchAttribute at[1];
                            at[0].id = cudaLaunchAttributeClusterDimension;
                            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                            cfg.attrs = at; cfg.numAttrs = 1;
                            cudaLaunchKernelEx(&cfg, rate_kernel<2>, N, iters, mode, stride, d);
                        }
                        cudaEventRecord(e1);
                        cudaError_t err = cudaDeviceSynchronize();
                        if (err != cudaSuccess) { printf("ctas=%d mode=%d N=%d grid=%d: %s\n", ctas, mode, N, grid, cudaGetErrorString(err)); return 1; }
                        float ms; cudaEventElapsedTime(&ms, e0, e1);
                        std::vector<long long> h(grid);
                        cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
                        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
                        const double cyc = (double)mx / iters;
                        const double ideal = (double)N / 2 / 1;      // per-SM cycles per MMA at the tensor peak (both modes: per SM)
                        const double flops = 2.0 * 128 * ctas * N * 16 * iters * (ctas == 1 ? grid : grid / 2);
                        printf("cta_group::%d %s sbo=%4d N=%3d grid=%3d: %7.1f cycles/MMA (ideal %5.1f, %4.1f%%)  %7.3f ms  %7.1f TFLOP/s\n",
                               ctas, mode ? "MN-major" : "K-major ", stride, N, grid, cyc, ideal, 100 * ideal / cyc, ms,
                               flops / ms / 1e9);
                    }
                }
            }
    return 0;
}
