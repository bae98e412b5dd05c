// How fast can tcgen05.mma be ISSUED?  One CTA; W issuer warps (one elected lane each) issue `ITER` MMAs (M = 128, N = n, K = 16,
// kind::f16, operands in shared memory or A in TMEM) back to back into disjoint accumulator columns, then commit; the cycles
// from the first issue to the completion of the last commit are reported per MMA, next to the tensor-pipe time N/2 cycles.
// The CNN's layer-1 (N = 96) and layer-2 (N = 16) MMAs are short: if the issue costs more than N/2 cycles the issuer, not the
// tensor pipe, sets the pace -- and whether several issuing warps scale decides the kernel's warp roles.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_bin/mma_issue_probe tools/mma_issue_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 240

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

template <int N, bool TS>
__global__ void __launch_bounds__(256, 1) probe(long long *cyc, int warps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[8];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp < warps) {
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t adesc = umma_desc(smem_u32(smem), 128 * 16, 128), bdesc = umma_desc(smem_u32(smem) + 16384, N * 16, 128);
        const uint32_t d = tmem + (uint32_t)(warp * 96), a_t = tmem + 448u;
        const uint32_t bar = smem_u32(&bars[warp]);
        __syncwarp();
        t0 = clock64();
        if (leader) {
#pragma unroll 8
            for (int i = 0; i < ITER; ++i) {
                if (TS)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(d), "r"(a_t + (uint32_t)((i & 3) * 8)), "l"(bdesc + (uint64_t)((i & 3) * 2 * N)), "r"(idesc), "r"((uint32_t)(i & 7)) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(d), "l"(adesc + (uint64_t)((i & 3) * 512)), "l"(bdesc + (uint64_t)((i & 3) * 2 * N)), "r"(idesc), "r"((uint32_t)(i & 7)) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
        t1 = clock64();
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0u) : "memory");
        t2 = clock64();
        if ((threadIdx.x & 31) == 0) { cyc[warp * 2] = t1 - t0; cyc[warp * 2 + 1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int N, bool TS> void run(long long *cyc) {
    cudaFuncSetAttribute(probe<N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int warps : {1, 2, 4}) {
        long long h[8] = {}, best_issue = 1ll << 60, best_done = 1ll << 60;
        for (int rep = 0; rep < 5; ++rep) {
            probe<N, TS><<<1, 256, 48 * 1024>>>(cyc, warps);
            if (cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(cudaGetLastError())); return; }
            long long mi = 0, md = 0;
            for (int w = 0; w < warps; ++w) { if (h[2 * w] > mi) mi = h[2 * w]; if (h[2 * w + 1] > md) md = h[2 * w + 1]; }
            if (md < best_done) { best_done = md; best_issue = mi; }
        }
        printf("N = %3d, A in %s, %d issuing warp(s): issue %6.1f cycles per MMA per warp, all complete after %6.1f cycles per MMA (all warps' MMAs; tensor pipe alone: %d)\n",
               N, TS ? "TMEM" : "smem", warps, (double)best_issue / ITER, (double)best_done / (ITER * warps), N / 2);
    }
}

int main() {
    long long *cyc;
    cudaMalloc(&cyc, 64);
    run<192, false>(cyc);
    run<96, false>(cyc);
    run<96, true>(cyc);
    run<16, true>(cyc);
    return 0;
}
