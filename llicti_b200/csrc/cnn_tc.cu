// tcgen05 (5th-gen tensor core) version of the per-band interpolator CNN
// (LLICTI_nets.py:721-753 layer 0, :695-712 grouped 1x1 layers, :822-825 get_params).
//
// The network is four independent 3-layer MLPs (sigma, mu, weights, coupling) sharing one
// im2col tile.  One CTA owns ONE sub-network of one band: its bf16 weights (<= 46 KB) are
// fetched once with a TMA bulk copy and stay in shared memory while the CTA walks over
// 128-position tiles:
//
//   im2col (integers, exact in bf16) -> smem A0
//   tcgen05.mma  D0[128 x NP] = A0[128 x K0p] * W0^T     (accumulator in TMEM)
//   epilogue     H = bf16(relu(D0 + b0))                  -> smem (K-major operand of layer 1)
//   tcgen05.mma  D1 = H * W1^T ; epilogue H = bf16(relu(D1 + b1))
//   tcgen05.mma  D2[128 x 16] = H * W2^T ; epilogue params = D2 + b2 -> global (fp32)
//
// The 4*chs-wide activations never touch HBM.  Two CTAs are resident per SM (104 KB smem,
// 256 TMEM columns each), so one CTA's epilogue overlaps the other's MMAs.
//
// Operands use the un-swizzled K-major canonical layout ("interleave"): 16-byte chunks of 8
// bf16 along K, element (row, k) at (k/8)*LBO + (row/8)*SBO + (row%8)*16 + (k%8)*2 with
// SBO = 128 B and LBO = rows*16 B, i.e. [k-chunk][row][16 B].  Epilogue thread t owns row t,
// so its 16-byte stores are contiguous across a warp (bank-conflict free) and need no swizzle.
//
// A position's outputs depend only on its own receptive field and the fixed instruction
// sequence (no split-K, no atomics), so compress and decompres compute identical parameters.
#include <cuda_bf16.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace llicti {

constexpr int TC_M = 128;        // positions per tile = UMMA_M
constexpr int TC_THREADS = 128;  // one thread per tile row / TMEM lane
constexpr int TC_TMEM_COLS = 256;

struct TcGeom {
    int Hs, Ws, P;          // plane size
    int total;              // n * P positions
    int ntiles;
    int K0, K0p;            // layer-0 depth and its padding to a multiple of 16
    int NP;                 // padded sub-network width (96 for 88, 64 for 60); also K of layers 1, 2
    int group_bytes;        // packed bytes of one sub-network (weights + biases)
    int off_w1, off_w2, off_bias;   // byte offsets inside the packed group
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a barrier that never completes (a programming error) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version (Blackwell)
    return d;                   // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}

// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

#define TMEM_LD_X16(taddr, r)                                                                           \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                              \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"      \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),  \
                   "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),          \
                   "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                                \
                 : "r"(taddr))
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// bias + ReLU + bf16 of one 16-column slab of the accumulator -> two 16-byte operand chunks.
__device__ __forceinline__ void relu_pack16(const uint32_t *r, const float *bias, uint8_t *dst_chunk0, uint32_t chunk_stride) {
    uint32_t w[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const float a = fmaxf(__uint_as_float(r[2 * e]) + bias[2 * e], 0.f);
        const float b = fmaxf(__uint_as_float(r[2 * e + 1]) + bias[2 * e + 1], 0.f);
        const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
        w[e] = *reinterpret_cast<const uint32_t *>(&p);
    }
    *reinterpret_cast<uint4 *>(dst_chunk0) = make_uint4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<uint4 *>(dst_chunk0 + chunk_stride) = make_uint4(w[4], w[5], w[6], w[7]);
}

__global__ void __launch_bounds__(TC_THREADS, 2)
cnn_tc_kernel(const int16_t *__restrict__ planes, TcGeom tg, TapTable taps, const uint8_t *__restrict__ packed,
              float *__restrict__ params) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int g = blockIdx.x & 3;                 // sub-network of this CTA
    const int tile0 = blockIdx.x >> 2;
    const int tile_stride = gridDim.x >> 2;

    // ---- shared memory carve-up -------------------------------------------------------------
    uint8_t *sW0 = smem;                                         // [K0p/8][NP][16 B]
    uint8_t *sW1 = smem + tg.off_w1;                             // [NP/8][NP][16 B]
    uint8_t *sW2 = smem + tg.off_w2;                             // [NP/8][16][16 B]
    const float *sBias = reinterpret_cast<const float *>(smem + tg.off_bias);   // b0[NP] b1[NP] b2[16]
    uint8_t *sA0 = smem + ((tg.group_bytes + 127) & ~127);       // [K0p/8][128][16 B]
    uint8_t *sH = sA0 + (tg.K0p / 8) * TC_M * 16;                // [NP/8][128][16 B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(sH + (tg.NP / 8) * TC_M * 16);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const uint32_t bar_w = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {     // weights + biases of this sub-network: one TMA bulk copy
        mbar_expect_tx(bar_w, (uint32_t)tg.group_bytes);
        tma_bulk_g2s(smem_u32(sW0), packed + (size_t)g * tg.group_bytes, (uint32_t)tg.group_bytes, bar_w);
    }
    mbar_wait(bar_w, 0);

    const uint32_t idescN = umma_idesc(tg.NP), idesc16 = umma_idesc(16);
    const uint32_t d0 = tmem, d1 = tmem + (uint32_t)tg.NP, d2 = tmem + 2u * (uint32_t)tg.NP;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;     // this warp's TMEM lanes
    const uint32_t a_lbo = TC_M * 16, w_lbo = (uint32_t)tg.NP * 16, w2_lbo = 16 * 16;
    const float *b0 = sBias, *b1 = sBias + tg.NP, *b2 = sBias + 2 * tg.NP;
    uint32_t phase = 0;

    for (int tile = tile0; tile < tg.ntiles; tile += tile_stride) {
        // ---- im2col with replicate padding: row `tid` of the tile ------------------------------
        const int q = tile * TC_M + tid;
        const bool valid = q < tg.total;
        const int qq = valid ? q : tg.total - 1;
        const int img = qq / tg.P, p = qq - img * tg.P;
        const int i = p / tg.Ws, j = p - i * tg.Ws;
        const int16_t *pl = planes + (size_t)img * 12 * tg.P;
        for (int kc = 0; kc < tg.K0p / 8; ++kc) {
            uint32_t w[4];
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                float v[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int k = kc * 8 + e2 * 2 + h;
                    float x = 0.f;
                    if (k < tg.K0) {
                        const int rr = min(max(i + taps.dy[k], 0), tg.Hs - 1);
                        const int cc = min(max(j + taps.dx[k], 0), tg.Ws - 1);
                        x = (float)pl[(size_t)(taps.phase[k] * 3 + taps.chan[k]) * tg.P + (size_t)rr * tg.Ws + cc];
                    }
                    v[h] = x;
                }
                const __nv_bfloat162 pk = __floats2bfloat162_rn(v[0], v[1]);   // |x| <= 255: exact
                w[e2] = *reinterpret_cast<const uint32_t *>(&pk);
            }
            *reinterpret_cast<uint4 *>(sA0 + (size_t)kc * a_lbo + tid * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- layer 0 ----------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            for (int ks = 0; ks < tg.K0p / 16; ++ks)
                umma_bf16(d0, umma_desc(smem_u32(sA0) + ks * 2 * a_lbo, a_lbo, 128),
                          umma_desc(smem_u32(sW0) + ks * 2 * w_lbo, w_lbo, 128), idescN, ks > 0);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase);
        phase ^= 1;
        tc_fence_after();
        for (int c16 = 0; c16 < tg.NP / 16; ++c16) {
            uint32_t r[16];
            TMEM_LD_X16(d0 + lane_base + c16 * 16, r);
            tmem_ld_wait();
            relu_pack16(r, b0 + c16 * 16, sH + (size_t)(c16 * 2) * a_lbo + tid * 16, a_lbo);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- layer 1 ----------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            for (int ks = 0; ks < tg.NP / 16; ++ks)
                umma_bf16(d1, umma_desc(smem_u32(sH) + ks * 2 * a_lbo, a_lbo, 128),
                          umma_desc(smem_u32(sW1) + ks * 2 * w_lbo, w_lbo, 128), idescN, ks > 0);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase);
        phase ^= 1;
        tc_fence_after();
        for (int c16 = 0; c16 < tg.NP / 16; ++c16) {
            uint32_t r[16];
            TMEM_LD_X16(d1 + lane_base + c16 * 16, r);
            tmem_ld_wait();
            relu_pack16(r, b1 + c16 * 16, sH + (size_t)(c16 * 2) * a_lbo + tid * 16, a_lbo);   // layer-1 MMAs are complete
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();

        // ---- layer 2 ----------------------------------------------------------------------------
        if (tid == 0) {
            tc_fence_after();
            for (int ks = 0; ks < tg.NP / 16; ++ks)
                umma_bf16(d2, umma_desc(smem_u32(sH) + ks * 2 * a_lbo, a_lbo, 128),
                          umma_desc(smem_u32(sW2) + ks * 2 * w2_lbo, w2_lbo, 128), idesc16, ks > 0);
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, phase);
        phase ^= 1;
        tc_fence_after();
        {
            uint32_t r[16];
            TMEM_LD_X16(d2 + lane_base, r);
            tmem_ld_wait();
            if (valid) {
                float *o = params + (size_t)img * kParamCh * tg.P + (size_t)(g * 15) * tg.P + p;
#pragma unroll
                for (int c = 0; c < 15; ++c) o[(size_t)c * tg.P] = __uint_as_float(r[c]) + b2[c];
            }
        }
        tc_fence_before();     // TMEM reads of this tile are ordered before the next tile's MMAs by the next barrier
    }

    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TC_TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Host side: packing and launch
// ---------------------------------------------------------------------------------------------
struct TcBand {
    uint8_t *packed = nullptr;   // device: 4 sub-networks back to back
    TcGeom g{};                  // shape-independent fields filled at pack time
    size_t smem_bytes = 0;
};
struct TcWeights {
    TcBand band[3];
};

static uint16_t f2bf(float f) {   // round to nearest even
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(r >> 16);
}

struct BranchDefTc { int band, kh, kw; };
static const BranchDefTc kBr[6] = {{0, 4, 4}, {1, 3, 4}, {1, 4, 3}, {2, 4, 3}, {2, 3, 4}, {2, 4, 4}};

int tc_pack_weights(llicti_ctx *ctx, const llicti_weights &w) {
    const int G = ctx->cfg.chs, Ch = 4 * G;
    const int NP = (G + 15) / 16 * 16;   // 88 -> 96, 60 -> 64
    TcWeights *tw = new TcWeights();
    for (int band = 0; band < 3; ++band) {
        const int K0 = ctx->taps[band].K0, K0p = (K0 + 15) / 16 * 16;
        const int w0_bytes = (K0p / 8) * NP * 16, w1_bytes = (NP / 8) * NP * 16, w2_bytes = (NP / 8) * 16 * 16;
        const int bias_bytes = (2 * NP + 16) * 4;
        const int group_bytes = w0_bytes + w1_bytes + w2_bytes + bias_bytes;
        std::vector<uint8_t> host((size_t)4 * group_bytes, 0);
        // layer-0 weights in the kernel's k order (branch, c, dy, dx), scaled by 1/255 (the kernel
        // feeds integer sample values, the reference feeds value/255)
        std::vector<float> w0((size_t)K0 * Ch, 0.f), b0(Ch, 0.f);
        int k = 0;
        for (int br = 0; br < 6; ++br) {
            if (kBr[br].band != band) continue;
            for (int c = 0; c < 3; ++c)
                for (int dy = 0; dy < kBr[br].kh; ++dy)
                    for (int dx = 0; dx < kBr[br].kw; ++dx, ++k)
                        for (int ch = 0; ch < Ch; ++ch)
                            w0[(size_t)k * Ch + ch] =
                                w.l0_w[br][(((size_t)ch * 3 + c) * kBr[br].kh + dy) * kBr[br].kw + dx] / 255.0f;
            for (int ch = 0; ch < Ch; ++ch) b0[ch] += w.l0_b[br][ch];
        }
        for (int g = 0; g < 4; ++g) {
            uint8_t *base = host.data() + (size_t)g * group_bytes;
            uint16_t *p0 = reinterpret_cast<uint16_t *>(base);
            for (int kk = 0; kk < K0; ++kk)
                for (int n = 0; n < G; ++n)
                    p0[((size_t)(kk / 8) * NP + n) * 8 + kk % 8] = f2bf(w0[(size_t)kk * Ch + g * G + n]);
            uint16_t *p1 = reinterpret_cast<uint16_t *>(base + w0_bytes);
            for (int in = 0; in < G; ++in)
                for (int n = 0; n < G; ++n)
                    p1[((size_t)(in / 8) * NP + n) * 8 + in % 8] = f2bf(w.l1_w[band][(size_t)(g * G + n) * G + in]);
            uint16_t *p2 = reinterpret_cast<uint16_t *>(base + w0_bytes + w1_bytes);
            for (int in = 0; in < G; ++in)
                for (int n = 0; n < 15; ++n)
                    p2[((size_t)(in / 8) * 16 + n) * 8 + in % 8] = f2bf(w.l2_w[band][(size_t)(g * 15 + n) * G + in]);
            float *pb = reinterpret_cast<float *>(base + w0_bytes + w1_bytes + w2_bytes);
            for (int n = 0; n < G; ++n) { pb[n] = b0[g * G + n]; pb[NP + n] = w.l1_b[band][g * G + n]; }
            for (int n = 0; n < 15; ++n) pb[2 * NP + n] = w.l2_b[band][g * 15 + n];
        }
        TcBand &tb = tw->band[band];
        LLICTI_CUDA(cudaMalloc((void **)&tb.packed, host.size()));
        LLICTI_CUDA(cudaMemcpy(tb.packed, host.data(), host.size(), cudaMemcpyHostToDevice));
        tb.g.K0 = K0; tb.g.K0p = K0p; tb.g.NP = NP; tb.g.group_bytes = group_bytes;
        tb.g.off_w1 = w0_bytes; tb.g.off_w2 = w0_bytes + w1_bytes; tb.g.off_bias = w0_bytes + w1_bytes + w2_bytes;
        tb.smem_bytes = (size_t)((group_bytes + 127) & ~127) + (size_t)(K0p / 8) * TC_M * 16 + (size_t)(NP / 8) * TC_M * 16 + 64;
    }
    ctx->tc_weights = tw;
    return LLICTI_OK;
}

void tc_free_weights(llicti_ctx *ctx) {
    TcWeights *tw = static_cast<TcWeights *>(ctx->tc_weights);
    if (!tw) return;
    for (auto &b : tw->band) cudaFree(b.packed);
    delete tw;
    ctx->tc_weights = nullptr;
}

int launch_cnn_tc(llicti_ctx *ctx, int band, const int16_t *planes, int n, int Hs, int Ws, float *params, cudaStream_t st) {
    ProfScope prof_(ctx, KC_CNN, st);
    TcWeights *tw = static_cast<TcWeights *>(ctx->tc_weights);
    LLICTI_REQUIRE(tw, "tcgen05 weights are not packed");
    TcBand &tb = tw->band[band];
    TcGeom tg = tb.g;
    tg.Hs = Hs; tg.Ws = Ws; tg.P = Hs * Ws;
    const long long total = (long long)n * tg.P;
    LLICTI_REQUIRE(total < (1ll << 31), "batch too large for one CNN launch");
    tg.total = (int)total;
    tg.ntiles = (tg.total + TC_M - 1) / TC_M;
    static int sm_count = 0;
    static size_t attr_smem = 0;
    if (!sm_count) {
        int dev = 0;
        LLICTI_CUDA(cudaGetDevice(&dev));
        LLICTI_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    if (tb.smem_bytes > attr_smem) {
        LLICTI_CUDA(cudaFuncSetAttribute(cnn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb.smem_bytes));
        attr_smem = tb.smem_bytes;
    }
    // persistent grid: 2 CTAs per SM, a multiple of 4 (one sub-network per CTA), no more than the work
    int ctas = std::min(2 * sm_count / 4 * 4, tg.ntiles * 4);
    cnn_tc_kernel<<<ctas, TC_THREADS, tb.smem_bytes, st>>>(planes, tg, ctx->taps[band], tb.packed, params);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
