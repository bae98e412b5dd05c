// Dependent-chain latency probe for the instructions on the serial coder chains (one warp, one SM).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/lat_probe tools/lat_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N 512
template <int OP>
__global__ void probe(uint32_t seed, uint32_t *out, long long *cyc) {
    uint32_t x = seed + threadIdx.x, y = seed * 3u + 1u, z = seed ^ 0x9e3779b9u;
    const int lane = threadIdx.x & 31;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = x + y;                                                 // IADD
        if (OP == 1) x = (uint32_t)(((uint64_t)x * y + y) >> 16);               // IMAD.WIDE + SHF.R.U64
        if (OP == 2) x = (uint32_t)__clz((int)(x | 1u)) + y;                    // FLO + IADD
        if (OP == 3) x = (uint32_t)__popc(x) + y;                               // POPC + IADD
        if (OP == 4) x = __ballot_sync(0xffffffffu, x > (uint32_t)lane) + y;    // ISETP + VOTE + IADD
        if (OP == 5) x = __shfl_sync(0xffffffffu, x, (int)(x & 31u)) + 1u;      // SHFL.IDX (dependent index) + IADD
        if (OP == 6) x = __shfl_sync(0xffffffffu, x + (uint32_t)lane, 7) + 1u;  // SHFL.IDX (fixed index)
        if (OP == 7) x = __reduce_max_sync(0xffffffffu, x ^ (uint32_t)lane) + y;        // REDUX
        if (OP == 8) x = __funnelshift_l(y, x, (x & 7u) + 1u);                  // SHF.L.W dependent shift
        if (OP == 9) {                                                           // ballot -> popc -> shfl (decoder selection)
            const uint32_t b = __ballot_sync(0xffffffffu, x >= (uint32_t)lane * 0x01000000u);
            const uint32_t li = (uint32_t)__popc(b) - 1u;
            x = __shfl_sync(0xffffffffu, x * 5u + (uint32_t)lane, (int)(li & 31u)) + z;
        }
        if (OP == 10) x = (x << (y & 31u)) & 0x7FFFFFFFu | (x >> 7);            // SHF + LOP3
        if (OP == 11) {                                                          // redux.max select then shfl
            const uint32_t li = __reduce_max_sync(0xffffffffu, x >= (uint32_t)lane * 0x01000000u ? (uint32_t)lane : 0u);
            x = __shfl_sync(0xffffffffu, x * 5u + (uint32_t)lane, (int)li) + z;
        }
        if (OP == 12) x = __fns(x | 1u, 0, 1) + y;                               // find-nth-set
        if (OP == 13) x = (uint32_t)__ffs((int)x) + y;                           // BREV+FLO
        if (OP == 14) x = (uint32_t)__float_as_int((float)(x | 1u)) + y;         // I2F
        if (OP == 15) x = (uint32_t)(((uint64_t)x * y) >> 32) + z;               // IMAD.HI
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; }
    out[threadIdx.x] = x;
}

template <int OP> void run(const char *name, uint32_t *out, long long *cyc) {
    long long h = 0, best = 1ll << 60;
    for (int r = 0; r < 5; ++r) {
        probe<OP><<<1, 32>>>(12345u + r, out, cyc);
        cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        if (h < best) best = h;
    }
    printf("%-48s %7.2f cycles per iteration\n", name, (double)best / N);
}

int main() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    run<0>("IADD", out, cyc);
    run<1>("IMAD.WIDE.U32(+c) + SHF.R.U64 16", out, cyc);
    run<2>("FLO(clz) + IADD", out, cyc);
    run<3>("POPC + IADD", out, cyc);
    run<4>("ISETP + VOTE.ballot + IADD", out, cyc);
    run<5>("SHFL.IDX dependent index + IADD", out, cyc);
    run<6>("IADD + SHFL.IDX fixed index + IADD", out, cyc);
    run<7>("LOP + REDUX.MAX + IADD", out, cyc);
    run<8>("LOP+IADD + SHF.L.W", out, cyc);
    run<9>("ballot -> popc -> shfl selection", out, cyc);
    run<10>("SHF + LOP3 chain", out, cyc);
    run<11>("redux.max -> shfl selection", out, cyc);
    run<12>("fns + IADD", out, cyc);
    run<13>("ffs + IADD", out, cyc);
    run<14>("I2F + IADD", out, cyc);
    run<15>("IMAD.HI + IADD", out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
