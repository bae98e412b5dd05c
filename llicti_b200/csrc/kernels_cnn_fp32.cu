// fp32 CUDA-core version of the per-band interpolator CNN (LLICTI_nets.py:721-753 layer 0
// with replicate padding, :695-712 grouped 1x1 layers, :822-825 get_params).
//
// This is the exactness reference of the repo: plain fp32 FMA chains in a fixed k order, so
// a position's 60 outputs depend only on its own receptive field -- identical whether the
// position is evaluated during compress or, band by band, during decompres.  The tcgen05
// kernel (cnn_tc.cu) is checked against it.
//
// Tiling: one CTA = 64 consecutive positions of one image; the four sub-networks (sigma, mu,
// weights, coupling) are independent after the shared im2col tile, so they are evaluated one
// after the other with 45 KB of activations in shared memory.
#include "common.cuh"

namespace llicti {

constexpr int TM = 64;        // positions per CTA
constexpr int NT = 128;       // threads per CTA
constexpr int PPT = 8;        // positions per thread (8 position groups x 16 channel groups)

template <int G>
struct CnnShape {
    static constexpr int CPT = (G == 88) ? 6 : 4;   // channels per thread -> 15 channel groups
    static constexpr int GP = G + 8;                // padded row stride of the weight matrices
    static_assert(15 * CPT >= G, "channel groups must cover the sub-network width");
};

template <int G>
__global__ void __launch_bounds__(NT)
cnn_fp32_kernel(const int16_t *__restrict__ planes, int Hs, int Ws, int K0, TapTable taps, BandWeightsF32 w,
                int div255_recip, float *__restrict__ params) {
    using S = CnnShape<G>;
    constexpr int CPT = S::CPT, GP = S::GP;
    extern __shared__ float smem[];
    float *A0 = smem;                  // [K0][TM]
    float *H1 = A0 + 128 * TM;         // [G][TM]
    float *H2 = H1 + G * TM;           // [G][TM]

    const int P = Hs * Ws;
    const int img = blockIdx.y;
    const int p0 = blockIdx.x * TM;
    const int16_t *pl = planes + (size_t)img * 12 * P;
    const int tid = threadIdx.x;

    // ---- im2col with replicate padding (clamped indices) --------------------------------
    {   // a thread stages one position's column: its row / column once, then every second tap
        const int q = tid & (TM - 1);
        const int p = min(p0 + q, P - 1);
        const int i = p / Ws, j = p - i * Ws;
#pragma unroll 4
        for (int k = tid / TM; k < K0; k += NT / TM) {
            const int rr = min(max(i + taps.dy[k], 0), Hs - 1);
            const int cc = min(max(j + taps.dx[k], 0), Ws - 1);
            const float v = (float)pl[(size_t)(taps.phase[k] * 3 + taps.chan[k]) * P + (size_t)rr * Ws + cc];
            A0[k * TM + q] = div255_recip ? __fmul_rn(v, 1.0f / 255.0f) : __fdiv_rn(v, 255.0f);
        }
    }
    __syncthreads();

    const int pg = tid & 7;        // position group: positions pg*8 .. pg*8+7
    const int cg = tid >> 3;       // channel group 0..15 (15 is idle in the wide layers)
    float *out_base = params + (size_t)img * kParamCh * P;

    for (int g = 0; g < 4; ++g) {
        // ---- layer 0: [TM x K0] * [K0 x G] + bias, ReLU -> H1 ----------------------------
        if (cg < 15) {
            float acc[CPT][PPT];
            const float *wp = w.w0 + (size_t)g * K0 * GP + cg * CPT;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int ch = cg * CPT + c;
                const float b = ch < G ? __ldg(w.b0 + g * G + ch) : 0.f;
#pragma unroll
                for (int q = 0; q < PPT; ++q) acc[c][q] = b;
            }
            for (int k = 0; k < K0; ++k) {
                const float4 a0 = *reinterpret_cast<const float4 *>(A0 + k * TM + pg * PPT);
                const float4 a1 = *reinterpret_cast<const float4 *>(A0 + k * TM + pg * PPT + 4);
                const float av[PPT] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float wv[CPT];             // one row segment of the packed weights: 8-byte (CPT = 6) / 16-byte (CPT = 4) aligned
                if constexpr (CPT == 6) {
                    const float2 w01 = __ldg(reinterpret_cast<const float2 *>(wp + (size_t)k * GP));
                    const float2 w23 = __ldg(reinterpret_cast<const float2 *>(wp + (size_t)k * GP + 2));
                    const float2 w45 = __ldg(reinterpret_cast<const float2 *>(wp + (size_t)k * GP + 4));
                    wv[0] = w01.x; wv[1] = w01.y; wv[2] = w23.x; wv[3] = w23.y; wv[4] = w45.x; wv[5] = w45.y;
                } else {
                    const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)k * GP));
                    wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
                }
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) acc[c][q] = fmaf(av[q], wv[c], acc[c][q]);
                }
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int ch = cg * CPT + c;
                if (ch < G) {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) H1[ch * TM + pg * PPT + q] = fmaxf(acc[c][q], 0.f);
                }
            }
        }
        __syncthreads();
        // ---- layer 1: [TM x G] * [G x G] + bias, ReLU -> H2 --------------------------------
        if (cg < 15) {
            float acc[CPT][PPT];
            const float *wp = w.w1 + (size_t)g * G * GP + cg * CPT;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int ch = cg * CPT + c;
                const float b = ch < G ? __ldg(w.b1 + g * G + ch) : 0.f;
#pragma unroll
                for (int q = 0; q < PPT; ++q) acc[c][q] = b;
            }
            for (int k = 0; k < G; ++k) {
                const float4 a0 = *reinterpret_cast<const float4 *>(H1 + k * TM + pg * PPT);
                const float4 a1 = *reinterpret_cast<const float4 *>(H1 + k * TM + pg * PPT + 4);
                const float av[PPT] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float wv[CPT];             // one row segment of the packed weights: 8-byte (CPT = 6) / 16-byte (CPT = 4) aligned
                if constexpr (CPT == 6) {
                    const float2 w01 = __ldg(reinterpret_cast<const float2 *>(wp + (size_t)k * GP));
                    const float2 w23 = __ldg(reinterpret_cast<const float2 *>(wp + (size_t)k * GP + 2));
                    const float2 w45 = __ldg(reinterpret_cast<const float2 *>(wp + (size_t)k * GP + 4));
                    wv[0] = w01.x; wv[1] = w01.y; wv[2] = w23.x; wv[3] = w23.y; wv[4] = w45.x; wv[5] = w45.y;
                } else {
                    const float4 w4 = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)k * GP));
                    wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
                }
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) acc[c][q] = fmaf(av[q], wv[c], acc[c][q]);
                }
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int ch = cg * CPT + c;
                if (ch < G) {
#pragma unroll
                    for (int q = 0; q < PPT; ++q) H2[ch * TM + pg * PPT + q] = fmaxf(acc[c][q], 0.f);
                }
            }
        }
        __syncthreads();
        // ---- layer 2: [TM x G] * [G x 15] + bias -> params ---------------------------------
        if (cg < 15) {
            float acc[PPT];
            const float b = __ldg(w.b2 + g * 15 + cg);
#pragma unroll
            for (int q = 0; q < PPT; ++q) acc[q] = b;
            const float *wp = w.w2 + (size_t)g * G * 16 + cg;
            for (int k = 0; k < G; ++k) {
                const float4 a0 = *reinterpret_cast<const float4 *>(H2 + k * TM + pg * PPT);
                const float4 a1 = *reinterpret_cast<const float4 *>(H2 + k * TM + pg * PPT + 4);
                const float av[PPT] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float wv = __ldg(wp + (size_t)k * 16);
#pragma unroll
                for (int q = 0; q < PPT; ++q) acc[q] = fmaf(av[q], wv, acc[q]);
            }
            float *o = out_base + (size_t)(g * 15 + cg) * P + p0 + pg * PPT;
#pragma unroll
            for (int q = 0; q < PPT; ++q)
                if (p0 + pg * PPT + q < P) o[q] = acc[q];
        }
        __syncthreads();
    }
}

template <int G>
static int launch_cnn_fp32_t(llicti_ctx *ctx, const TapTable &t, int band, const int16_t *planes, dim3 grid, size_t smem, int Hs, int Ws,
                             float *params, cudaStream_t st) {
    // (the opt-in is a per-device attribute of the function: set at every launch, a host-side call of about a microsecond)
    LLICTI_CUDA(cudaFuncSetAttribute(cnn_fp32_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cnn_fp32_kernel<G><<<grid, NT, smem, st>>>(planes, Hs, Ws, t.K0, t, ctx->wf32[band], ctx->num.div255_recip, params);
    return LLICTI_OK;
}

int launch_cnn_fp32(llicti_ctx *ctx, int band, const int16_t *planes, int n, int Hs, int Ws, float *params,
                    cudaStream_t st) {
    ProfScope prof_(ctx, KC_CNN, st);
    const int G = ctx->cfg.chs;
    const int P = Hs * Ws;
    const TapTable &t = ctx->taps[band];
    const size_t smem = (size_t)(128 * TM + 2 * G * TM) * sizeof(float);
    dim3 grid((P + TM - 1) / TM, n);
    int rc;
    if (G == 88) {
        rc = launch_cnn_fp32_t<88>(ctx, t, band, planes, grid, smem, Hs, Ws, params, st);
    } else if (G == 60) {
        rc = launch_cnn_fp32_t<60>(ctx, t, band, planes, grid, smem, Hs, Ws, params, st);
    } else {
        set_error("cnn: unsupported sub-network width %d (88 or 60)", G);
        return LLICTI_E_ARG;
    }
    if (rc) return rc;
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
