// TMEM read-out rate probe: W warps of one CTA read their lane quadrant of TMEM with back-to-back tcgen05.ld.32x32b.x32
// (4 KB per warp instruction) and the cycles per byte are reported for W = 1, 2, 4, 8, 16.  Answers whether the read
// path is 64 B/clk per SM or per sub-partition (the CNN's epilogues read 212 KB per tile and CTA).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 256
#define TMEM_LD_X32(taddr, r)                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                              \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "             \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),  \
                   "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),          \
                   "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),       \
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),       \
                   "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),       \
                   "=r"(r[31])                                                                          \
                 : "r"(taddr))

// MODE 0: loads only, one wait at the end of each group of 3; MODE 1: loads + the CNN epilogue's conversion (cvt.rn.relu.bf16x2 +
// 16-byte shared stores), to see what the conversion adds.
template <int MODE>
__global__ void __launch_bounds__(512, 1) probe(uint32_t *out, long long *cyc) {
    __shared__ uint32_t slot;
    extern __shared__ __align__(16) uint8_t hbuf[];      // [12 chunks][512 threads][16 B]
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    uint32_t r[96];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
        const uint32_t col = (uint32_t)((i & 3) * 96);
        TMEM_LD_X32(tmem + col, r);
        TMEM_LD_X32(tmem + col + 32, (r + 32));
        TMEM_LD_X32(tmem + col + 64, (r + 64));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 96; k += 32) acc ^= r[k];
        } else {
            uint8_t *h = hbuf + threadIdx.x * 16;
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                uint32_t w[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w[e]) : "f"(__uint_as_float(r[c * 8 + 2 * e + 1])), "f"(__uint_as_float(r[c * 8 + 2 * e])));
                *reinterpret_cast<uint4 *>(h + (size_t)c * 512 * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
    out[threadIdx.x] = acc + hbuf[threadIdx.x];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 16 * 12);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * 16 * 12);
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {1, 2, 4, 8, 16}) {
            long long h = 0, best = 1ll << 60;
            for (int rep = 0; rep < 5; ++rep) {
                if (mode == 0) probe<0><<<1, warps * 32, 512 * 16 * 12>>>(out, cyc); else probe<1><<<1, warps * 32, 512 * 16 * 12>>>(out, cyc);
                if (cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                if (h < best) best = h;
            }
            const double bytes = (double)warps * ITER * 3 * 4096;
            printf("mode %d (%s) warps %2d: %8lld cycles, %7.1f B/clk per SM, %6.1f B/clk per warp\n", mode, mode ? "ld+cvt+sts" : "ld only", warps,
                   best, bytes / best, bytes / best / warps);
        }
    return 0;
}
