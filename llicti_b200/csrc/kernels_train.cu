// Backward of the rate-estimation path: gradients of a loss over LLICTI.forward's self-informations with respect to
// every weight of the interpolator CNNs -- the device half of the reference's training step
// (agents/llicti_agent.py:48-83: `self.model(x)` -> TrainRLossList -> `.backward()`; the autograd graph it
// differentiates is LLICTI_nets.py:101-123, 827-935 and entropy_layer_nets.py:121-139, 160-183, with
// compressai's LowerBound on the spreads, the mixture weights and the likelihood).
//
// Three kernels per (scale, band):
//   cnn_forward_train_kernel<G, K0>  (forward pass) the band's 60 network outputs from the fp32 planes, kept for the backward pass
//   self_info_grad_kernel            one thread per position: d loss / d (60 network outputs), in place over the outputs
//   cnn_backward_kernel<G, K0>       persistent CTAs, one of the four sub-networks each (blockIdx.y), its packed weights
//                                    resident in shared memory: a tile of 64 positions is re-staged (im2col), the two hidden
//                                    layers are recomputed in shared memory (nothing but the 60 outputs ever went to HBM),
//                                    and the five products of the backward pass run on the tile; the CTA's weight gradients
//                                    accumulate in REGISTERS over all its tiles and leave with one atomicAdd per element.
// fp32 throughout (the reference trains in fp32).  Results equal torch.autograd's up to summation order.
#include "common.cuh"
#include "gmm.cuh"

namespace llicti {

// ---- likelihood backward ---------------------------------------------------------------------------
// compressai.ops.LowerBound: y = max(x, bound); the gradient passes where x >= bound or where it would raise x.
__device__ __forceinline__ float lb_grad(float x, float bound, float g) { return (x >= bound || g < 0.f) ? g : 0.f; }

__global__ void __launch_bounds__(128)
self_info_grad_kernel(float *__restrict__ params, const float *__restrict__ fplanes, const float *__restrict__ gsinfo, int band,
                      int P, NumericsProfile np) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (i >= P) return;
    float *pp = params + (size_t)img * kParamCh * P + i;
    const float *yt = fplanes + (size_t)img * 12 * P + (size_t)(3 * (band + 1)) * P + i;
    const float *gs = gsinfo + (size_t)img * 9 * P + (size_t)(3 * band) * P + i;
    const float y0 = yt[0], y1 = yt[P], y2 = yt[2 * (size_t)P];
    const float half = (float)(0.5 / 255.0), sb = (float)(0.11 / 255.0), wb = 1e-6f, lkb = 1e-9f;
    const float v[3] = {y0, y1, y2};
    const float kInvSqrt2Pi = 0.3989422804014327f, kInvLn2 = 1.4426950408889634f;
#pragma unroll
    for (int clr = 0; clr < 3; ++clr) {
        float sraw[kM], sg[kM], mu[kM], wraw[kM], w[kM], lm[kM], zu[kM], zl[kM], cpl0[kM], cpl1[kM];
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            sraw[m] = pp[(size_t)(clr * kM + m) * P];
            sg[m] = fmaxf(sraw[m], sb);
            mu[m] = pp[(size_t)((3 + clr) * kM + m) * P];
            wraw[m] = pp[(size_t)((6 + clr) * kM + m) * P];
            w[m] = fmaxf(wraw[m], wb);
            cpl0[m] = cpl1[m] = 0.f;
            if (clr == 1) { cpl0[m] = pp[(size_t)(9 * kM + m) * P]; mu[m] = __fadd_rn(mu[m], __fmul_rn(cpl0[m], y0)); }
            if (clr == 2) {
                cpl0[m] = pp[(size_t)(10 * kM + m) * P];
                cpl1[m] = pp[(size_t)(11 * kM + m) * P];
                mu[m] = __fadd_rn(mu[m], __fadd_rn(__fmul_rn(cpl0[m], y0), __fmul_rn(cpl1[m], y1)));
            }
        }
        const float den = sum5(w, np);
        float t[kM];
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            const float a = fabsf(__fsub_rn(v[clr], mu[m]));
            zu[m] = __fdiv_rn(__fsub_rn(half, a), sg[m]);
            zl[m] = __fdiv_rn(__fsub_rn(-half, a), sg[m]);
            const float up = __fmul_rn(0.5f, erfcf(__fmul_rn(-0.70710678118654752440f, zu[m])));
            const float lo = __fmul_rn(0.5f, erfcf(__fmul_rn(-0.70710678118654752440f, zl[m])));
            lm[m] = __fsub_rn(up, lo);
            t[m] = __fmul_rn(__fdiv_rn(w[m], den), lm[m]);
        }
        const float lik_raw = sum5(t, np);
        const float lik = fmaxf(lik_raw, lkb);
        // s = -log2(lik)
        const float g = gs[(size_t)clr * P];
        const float dlik = lb_grad(lik_raw, lkb, -g * kInvLn2 / lik);
        float dot = 0.f;                                    // sum_k dpi_k pi_k
#pragma unroll
        for (int m = 0; m < kM; ++m) dot += dlik * lm[m] * (w[m] / den);
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            const float pi = w[m] / den;
            const float dw = lb_grad(wraw[m], wb, (dlik * lm[m] - dot) / den);
            const float dl = dlik * pi;
            const float dzu = dl * kInvSqrt2Pi * __expf(-0.5f * zu[m] * zu[m]);
            const float dzl = -dl * kInvSqrt2Pi * __expf(-0.5f * zl[m] * zl[m]);
            const float dval = -(dzu + dzl) / sg[m];
            const float dsg = lb_grad(sraw[m], sb, -(dzu * zu[m] + dzl * zl[m]) / sg[m]);
            const float diff = __fsub_rn(v[clr], mu[m]);
            const float dmu = -dval * (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
            pp[(size_t)(clr * kM + m) * P] = dsg;
            pp[(size_t)((3 + clr) * kM + m) * P] = dmu;
            pp[(size_t)((6 + clr) * kM + m) * P] = dw;
            if (clr == 1) pp[(size_t)(9 * kM + m) * P] = dmu * y0;
            if (clr == 2) {
                pp[(size_t)(10 * kM + m) * P] = dmu * y0;
                pp[(size_t)(11 * kM + m) * P] = dmu * y1;
            }
        }
    }
}

int launch_self_info_grad(llicti_ctx *ctx, float *params, const float *fplanes, const float *gsinfo, int band, int n, int P,
                          cudaStream_t st) {
    ProfScope prof_(ctx, KC_BOUNDS, st);
    dim3 grid((P + 127) / 128, n);
    self_info_grad_kernel<<<grid, 128, 0, st>>>(params, fplanes, gsinfo, band, P, ctx->num);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

// ---- CNN backward -------------------------------------------------------------------------------------
constexpr int BT = 64;         // positions per tile
constexpr int BLD = BT + 4;    // row stride of the tile matrices: rows r and r+1 start four banks apart
constexpr int BNT = 256;       // threads per CTA: 16 position groups of 4 x 16 channel groups

template <int G>
struct BwdShape {
    static constexpr int CPT = (G == 88) ? 6 : 4;   // channels per thread in the layer products: 15 active channel groups
    static constexpr int GP = G + 8;                // row stride of the packed weights (pack_weights, api.cu)
    static_assert(15 * CPT >= G && G % 4 == 0, "channel groups must cover the sub-network width");
};

// Y[ch][q] = relu(b[ch] + sum_k X[k][q] * Wt[k][ch])   (X: [K][BLD], Wt: packed [K][GP], both in shared memory)
template <int G>
__device__ __forceinline__ void layer_forward(const float *X, int K, const float *Wt, const float *__restrict__ b,
                                              float *Y, int tid) {
    constexpr int CPT = BwdShape<G>::CPT, GP = BwdShape<G>::GP;
    const int pg = tid & 15, cg = tid >> 4;
    if (cg >= 15) return;
    float acc[CPT][4];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int ch = cg * CPT + c;
        const float bv = ch < G ? __ldg(b + ch) : 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[c][q] = bv;
    }
    const float *wp = Wt + cg * CPT;       // 8-byte (CPT = 6, GP = 96) / 16-byte (CPT = 4, GP = 68) aligned in every row
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
        const float4 a = *reinterpret_cast<const float4 *>(X + k * BLD + pg * 4);
        float wv[CPT];                     // (columns G .. GP-1 of the packed rows are zero padding)
        if constexpr (CPT == 6) {
            const float2 w01 = *reinterpret_cast<const float2 *>(wp + k * GP);
            const float2 w23 = *reinterpret_cast<const float2 *>(wp + k * GP + 2);
            const float2 w45 = *reinterpret_cast<const float2 *>(wp + k * GP + 4);
            wv[0] = w01.x; wv[1] = w01.y; wv[2] = w23.x; wv[3] = w23.y; wv[4] = w45.x; wv[5] = w45.y;
        } else {
            const float4 w4 = *reinterpret_cast<const float4 *>(wp + k * GP);
            wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            acc[c][0] = fmaf(a.x, wv[c], acc[c][0]);
            acc[c][1] = fmaf(a.y, wv[c], acc[c][1]);
            acc[c][2] = fmaf(a.z, wv[c], acc[c][2]);
            acc[c][3] = fmaf(a.w, wv[c], acc[c][3]);
        }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int ch = cg * CPT + c;
        if (ch < G)
            *reinterpret_cast<float4 *>(Y + ch * BLD + pg * 4) =
                make_float4(fmaxf(acc[c][0], 0.f), fmaxf(acc[c][1], 0.f), fmaxf(acc[c][2], 0.f), fmaxf(acc[c][3], 0.f));
    }
}

// dX[k][q] = (X[k][q] > 0) * sum_o Wt[k][o] * dY[o][q]   for k < G; o < No.  OUT may alias neither X nor dY.
template <int G>
__device__ __forceinline__ void layer_backward_data(const float *X, const float *Wt, int wstride, const float *dY, int No,
                                                    float *OUT, int tid) {
    constexpr int CPT = BwdShape<G>::CPT;
    const int pg = tid & 15, cg = tid >> 4;
    if (cg >= 15) return;
    float acc[CPT][4];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[c][q] = 0.f;
    for (int o = 0; o < No; o += 4) {       // No % 4 == 0; rows of Wt are 16-byte aligned at multiples of four columns
        const float4 d0 = *reinterpret_cast<const float4 *>(dY + (o + 0) * BLD + pg * 4);
        const float4 d1 = *reinterpret_cast<const float4 *>(dY + (o + 1) * BLD + pg * 4);
        const float4 d2 = *reinterpret_cast<const float4 *>(dY + (o + 2) * BLD + pg * 4);
        const float4 d3 = *reinterpret_cast<const float4 *>(dY + (o + 3) * BLD + pg * 4);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int k = min(cg * CPT + c, G - 1);
            const float4 wv = *reinterpret_cast<const float4 *>(Wt + k * wstride + o);
            acc[c][0] = fmaf(d3.x, wv.w, fmaf(d2.x, wv.z, fmaf(d1.x, wv.y, fmaf(d0.x, wv.x, acc[c][0]))));
            acc[c][1] = fmaf(d3.y, wv.w, fmaf(d2.y, wv.z, fmaf(d1.y, wv.y, fmaf(d0.y, wv.x, acc[c][1]))));
            acc[c][2] = fmaf(d3.z, wv.w, fmaf(d2.z, wv.z, fmaf(d1.z, wv.y, fmaf(d0.z, wv.x, acc[c][2]))));
            acc[c][3] = fmaf(d3.w, wv.w, fmaf(d2.w, wv.z, fmaf(d1.w, wv.y, fmaf(d0.w, wv.x, acc[c][3]))));
        }
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int k = cg * CPT + c;
        if (k < G) {
            const float4 x = *reinterpret_cast<const float4 *>(X + k * BLD + pg * 4);
            *reinterpret_cast<float4 *>(OUT + k * BLD + pg * 4) =
                make_float4(x.x > 0.f ? acc[c][0] : 0.f, x.y > 0.f ? acc[c][1] : 0.f, x.z > 0.f ? acc[c][2] : 0.f,
                            x.w > 0.f ? acc[c][3] : 0.f);
        }
    }
}

// acc[i][j] += sum_q A[i][q] * B[j][q]  (i < RA, j < RB; both multiples of 4).  A thread owns NB blocks of 4 x 4 outputs IN
// REGISTERS for the whole kernel (block tid + i * BNT); a block's rows are nbA / nbB apart, so that the lanes of a warp read
// CONSECUTIVE rows of B (four banks apart: conflict-free 16-byte loads) and mostly the same rows of A (broadcast).
template <int RA, int RB>
struct WGrad {
    static constexpr int nbA = RA / 4, nbB = RB / 4, NB = (nbA * nbB + BNT - 1) / BNT;
    static_assert(RA % 4 == 0 && RB % 4 == 0, "4 x 4 blocks");

    static __device__ __forceinline__ void zero(float (&acc)[NB][4][4]) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][r][c] = 0.f;
    }

    static __device__ __forceinline__ void accumulate(const float *A, const float *B, float (&acc)[NB][4][4], int tid) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int blk = tid + i * BNT;
            if (blk >= nbA * nbB) break;
            const int ib = blk / nbB, jb = blk - ib * nbB;
#pragma unroll 2
            for (int q = 0; q < BT; q += 4) {
                float4 a[4], b[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4 *>(A + (ib + r * nbA) * BLD + q);
#pragma unroll
                for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4 *>(B + (jb + c * nbB) * BLD + q);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        acc[i][r][c] = fmaf(a[r].x, b[c].x, fmaf(a[r].y, b[c].y, fmaf(a[r].z, b[c].z, fmaf(a[r].w, b[c].w, acc[i][r][c]))));
            }
        }
    }

    // one atomicAdd per element into the packed global accumulator (row stride dst_stride)
    static __device__ __forceinline__ void flush(const float (&acc)[NB][4][4], float *dst, int dst_stride, int tid) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int blk = tid + i * BNT;
            if (blk >= nbA * nbB) break;
            const int ib = blk / nbB, jb = blk - ib * nbB;
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) atomicAdd(dst + (size_t)(ib + r * nbA) * dst_stride + jb + c * nbB, acc[i][r][c]);
        }
    }
};

// ACC[j] += sum_q B[j][q]
__device__ __forceinline__ void bias_grad(const float *B, int Rb, float *ACC, int tid) {
    for (int j = tid; j < Rb; j += BNT) {
        float s = 0.f;
#pragma unroll 4
        for (int q = 0; q < BT; q += 4) {
            const float4 b = *reinterpret_cast<const float4 *>(B + j * BLD + q);
            s += (b.x + b.y) + (b.z + b.w);
        }
        ACC[j] += s;
    }
}

struct BandGradsF32 {      // packed like BandWeightsF32
    float *w0, *b0, *w1, *b1, *w2, *b2;
};

template <int G, int K0>
__global__ void __launch_bounds__(BNT, 1)
cnn_backward_kernel(const float *__restrict__ fplanes, int Hs, int Ws, TapTable taps, BandWeightsF32 w,
                    const float *__restrict__ dparams, int n, BandGradsF32 gr) {
    constexpr int GP = BwdShape<G>::GP;
    extern __shared__ __align__(16) float smem[];
    float *A0 = smem;                    // [K0][BLD]  staged receptive fields
    float *H1 = A0 + K0 * BLD;           // [G][BLD]
    float *H2 = H1 + G * BLD;            // [G][BLD]   later d H1
    float *D2 = H2 + G * BLD;            // [G][BLD]   d H2
    float *DO = D2 + G * BLD;            // [16][BLD]  d outputs of this sub-network (row 15: zero)
    float *W0 = DO + 16 * BLD;           // [K0][GP]   this sub-network's packed weights, resident for the whole kernel
    float *W1 = W0 + K0 * GP;            // [G][GP]
    float *W2 = W1 + G * GP;             // [G][16]
    float *GB0 = W2 + G * 16;            // [G]        bias gradients
    float *GB1 = GB0 + G;                // [G]
    float *GB2 = GB1 + G;                // [16]

    const int tid = threadIdx.x;
    const int g = blockIdx.y;
    const int P = Hs * Ws;
    const int tiles_per_img = (P + BT - 1) / BT;
    const long long tiles = (long long)tiles_per_img * n;
    if (tiles <= (long long)blockIdx.x) return;            // (whole CTA) no tile: nothing to add

    {   // weights in (16-byte copies: all three packed blocks are multiples of four floats and 16-byte aligned)
        const float4 *s0 = reinterpret_cast<const float4 *>(w.w0 + (size_t)g * K0 * GP);
        const float4 *s1 = reinterpret_cast<const float4 *>(w.w1 + (size_t)g * G * GP);
        const float4 *s2 = reinterpret_cast<const float4 *>(w.w2 + (size_t)g * G * 16);
        for (int e = tid; e < K0 * GP / 4; e += BNT) reinterpret_cast<float4 *>(W0)[e] = __ldg(s0 + e);
        for (int e = tid; e < G * GP / 4; e += BNT) reinterpret_cast<float4 *>(W1)[e] = __ldg(s1 + e);
        for (int e = tid; e < G * 16 / 4; e += BNT) reinterpret_cast<float4 *>(W2)[e] = __ldg(s2 + e);
    }
    for (int e = tid; e < 2 * G + 16; e += BNT) GB0[e] = 0.f;
    for (int e = tid; e < BLD; e += BNT) DO[15 * BLD + e] = 0.f;
    const float *b0 = w.b0 + g * G, *b1 = w.b1 + g * G;

    using WG2 = WGrad<G, 16>;
    using WG1 = WGrad<G, G>;
    using WG0 = WGrad<K0, G>;
    float gw2[WG2::NB][4][4], gw1[WG1::NB][4][4], gw0[WG0::NB][4][4];      // the CTA's weight gradients: registers
    WG2::zero(gw2);
    WG1::zero(gw1);
    WG0::zero(gw0);

    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int img = (int)(t / tiles_per_img);
        const int p0 = (int)(t - (long long)img * tiles_per_img) * BT;
        const float *pl = fplanes + (size_t)img * 12 * P;     // the float lifting's values, as the reference's graph sees them
        __syncthreads();                 // the previous tile's products have read A0 / H2 / D2 / DO (first tile: weights are in)
        {   // a thread stages ONE position's column: its row / column once, then every fourth tap and output row
            const int q = tid & (BT - 1), k4 = tid >> 6;
            const int p = min(p0 + q, P - 1);
            const int i = p / Ws, j = p - i * Ws;
#pragma unroll 4
            for (int k = k4; k < K0; k += BNT / BT) {
                const int rr = min(max(i + taps.dy[k], 0), Hs - 1);
                const int cc = min(max(j + taps.dx[k], 0), Ws - 1);
                A0[k * BLD + q] = pl[(size_t)(taps.phase[k] * 3 + taps.chan[k]) * P + (size_t)rr * Ws + cc];
            }
            const bool inside = p0 + q < P;                      // positions past the end: no gradient
            const float *dp = dparams + ((size_t)img * kParamCh + g * 15) * P + p;
#pragma unroll
            for (int o = k4; o < 15; o += BNT / BT) DO[o * BLD + q] = inside ? dp[(size_t)o * P] : 0.f;
        }
        __syncthreads();
        layer_forward<G>(A0, K0, W0, b0, H1, tid);
        __syncthreads();
        layer_forward<G>(H1, G, W1, b1, H2, tid);
        __syncthreads();
        // layer 2: d W2 = H2 dO^T, d b2, d H2 = relu'(H2) * (W2 dO)
        WG2::accumulate(H2, DO, gw2, tid);
        bias_grad(DO, 16, GB2, tid);
        layer_backward_data<G>(H2, W2, 16, DO, 16, D2, tid);      // (column 15 of the packed W2 and row 15 of dO are zero)
        __syncthreads();
        // layer 1: d W1 = H1 dH2^T, d b1, d H1 = relu'(H1) * (W1 dH2) -> H2's buffer
        WG1::accumulate(H1, D2, gw1, tid);
        bias_grad(D2, G, GB1, tid);
        layer_backward_data<G>(H1, W1, GP, D2, G, H2, tid);
        __syncthreads();
        // layer 0: d W0 = A0 dH1^T, d b0
        WG0::accumulate(A0, H2, gw0, tid);
        bias_grad(H2, G, GB0, tid);
    }
    WG0::flush(gw0, gr.w0 + (size_t)g * K0 * GP, GP, tid);
    WG1::flush(gw1, gr.w1 + (size_t)g * G * GP, GP, tid);
    WG2::flush(gw2, gr.w2 + (size_t)g * G * 16, 16, tid);
    __syncthreads();
    for (int e = tid; e < G; e += BNT) {
        atomicAdd(gr.b0 + g * G + e, GB0[e]);
        atomicAdd(gr.b1 + g * G + e, GB1[e]);
    }
    for (int e = tid; e < 15; e += BNT) atomicAdd(gr.b2 + g * 15 + e, GB2[e]);
}

// The fp32 CNN over the FLOAT planes (LLICTI.forward's float lifting: the values the reference's training / validation graph
// feeds its convolutions, an ulp away from integer / 255) with the backward kernel's organisation: persistent CTAs, one
// sub-network each, its packed weights resident in shared memory; a tile's two hidden layers and its 15 outputs.  Every
// output is the same fmaf chain (bias first, k ascending) as in cnn_fp32_kernel.  Serves forward() and the training step of
// fp32 contexts.
template <int G, int K0>
__global__ void __launch_bounds__(BNT, 1)
cnn_forward_train_kernel(const float *__restrict__ fplanes, int Hs, int Ws, TapTable taps, BandWeightsF32 w, int n,
                         float *__restrict__ params) {
    constexpr int GP = BwdShape<G>::GP;
    extern __shared__ __align__(16) float smem[];
    float *A0 = smem;                    // [K0][BLD]
    float *H1 = A0 + K0 * BLD;           // [G][BLD]
    float *H2 = H1 + G * BLD;            // [G][BLD]
    float *W0 = H2 + G * BLD;            // [K0][GP]
    float *W1 = W0 + K0 * GP;            // [G][GP]
    float *W2 = W1 + G * GP;             // [G][16]
    const int tid = threadIdx.x;
    const int g = blockIdx.y;
    const int P = Hs * Ws;
    const int tiles_per_img = (P + BT - 1) / BT;
    const long long tiles = (long long)tiles_per_img * n;
    if (tiles <= (long long)blockIdx.x) return;
    {
        const float4 *s0 = reinterpret_cast<const float4 *>(w.w0 + (size_t)g * K0 * GP);
        const float4 *s1 = reinterpret_cast<const float4 *>(w.w1 + (size_t)g * G * GP);
        const float4 *s2 = reinterpret_cast<const float4 *>(w.w2 + (size_t)g * G * 16);
        for (int e = tid; e < K0 * GP / 4; e += BNT) reinterpret_cast<float4 *>(W0)[e] = __ldg(s0 + e);
        for (int e = tid; e < G * GP / 4; e += BNT) reinterpret_cast<float4 *>(W1)[e] = __ldg(s1 + e);
        for (int e = tid; e < G * 16 / 4; e += BNT) reinterpret_cast<float4 *>(W2)[e] = __ldg(s2 + e);
    }
    const float *b0 = w.b0 + g * G, *b1 = w.b1 + g * G;
    const int pg = tid & 15, og = tid >> 4;             // layer 2: four positions x one output per thread (og < 15)
    const float b2 = og < 15 ? __ldg(w.b2 + g * 15 + og) : 0.f;

    for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int img = (int)(t / tiles_per_img);
        const int p0 = (int)(t - (long long)img * tiles_per_img) * BT;
        const float *pl = fplanes + (size_t)img * 12 * P;
        __syncthreads();                 // the previous tile's layers have read A0 / H1 / H2 (first tile: weights are in)
        {
            const int q = tid & (BT - 1), k4 = tid >> 6;
            const int p = min(p0 + q, P - 1);
            const int i = p / Ws, j = p - i * Ws;
#pragma unroll 4
            for (int k = k4; k < K0; k += BNT / BT) {
                const int rr = min(max(i + taps.dy[k], 0), Hs - 1);
                const int cc = min(max(j + taps.dx[k], 0), Ws - 1);
                A0[k * BLD + q] = pl[(size_t)(taps.phase[k] * 3 + taps.chan[k]) * P + (size_t)rr * Ws + cc];
            }
        }
        __syncthreads();
        layer_forward<G>(A0, K0, W0, b0, H1, tid);
        __syncthreads();
        layer_forward<G>(H1, G, W1, b1, H2, tid);
        __syncthreads();
        if (og < 15) {
            float acc[4] = {b2, b2, b2, b2};
#pragma unroll 4
            for (int k = 0; k < G; ++k) {
                const float4 a = *reinterpret_cast<const float4 *>(H2 + k * BLD + pg * 4);
                const float wv = W2[k * 16 + og];
                acc[0] = fmaf(a.x, wv, acc[0]);
                acc[1] = fmaf(a.y, wv, acc[1]);
                acc[2] = fmaf(a.z, wv, acc[2]);
                acc[3] = fmaf(a.w, wv, acc[3]);
            }
            float *o = params + ((size_t)img * kParamCh + g * 15 + og) * P + p0 + pg * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (p0 + pg * 4 + q < P) o[q] = acc[q];
        }
    }
}

// Packed gradient accumulators of the three bands: one allocation, laid out like the packed weights.
struct TrainState {
    float *base = nullptr;
    size_t floats = 0;
    BandGradsF32 g[3];
};

static size_t band_pack_floats(int G, int K0) { return (size_t)4 * K0 * (G + 8) + 4 * G + (size_t)4 * G * (G + 8) + 4 * G + (size_t)4 * G * 16 + 64; }

static int train_state(llicti_ctx *ctx, TrainState **out) {
    if (!ctx->train_state) {
        TrainState *ts = new TrainState();
        const int G = ctx->cfg.chs;
        for (int b = 0; b < 3; ++b) ts->floats += band_pack_floats(G, ctx->taps[b].K0);
        cudaError_t e = cudaMalloc((void **)&ts->base, ts->floats * sizeof(float));
        if (e != cudaSuccess) { delete ts; set_error("cudaMalloc(gradient accumulators): %s", cudaGetErrorString(e)); return LLICTI_E_CUDA; }
        float *p = ts->base;
        for (int b = 0; b < 3; ++b) {
            const int K0 = ctx->taps[b].K0, GP = G + 8;
            ts->g[b].w0 = p; p += (size_t)4 * K0 * GP;
            ts->g[b].b0 = p; p += 4 * G;
            ts->g[b].w1 = p; p += (size_t)4 * G * GP;
            ts->g[b].b1 = p; p += 4 * G;
            ts->g[b].w2 = p; p += (size_t)4 * G * 16;
            ts->g[b].b2 = p; p += 64;
        }
        ctx->train_state = ts;
    }
    *out = (TrainState *)ctx->train_state;
    return LLICTI_OK;
}

void train_free(llicti_ctx *ctx) {
    TrainState *ts = (TrainState *)ctx->train_state;
    if (!ts) return;
    cudaFree(ts->base);
    delete ts;
    ctx->train_state = nullptr;
}

int launch_train_zero_grads(llicti_ctx *ctx, cudaStream_t st) {
    TrainState *ts;
    int rc = train_state(ctx, &ts);
    if (rc) return rc;
    LLICTI_CUDA(cudaMemsetAsync(ts->base, 0, ts->floats * sizeof(float), st));
    return LLICTI_OK;
}

template <int G, int K0>
static int launch_backward_t(llicti_ctx *ctx, TrainState *ts, int band, const float *fplanes, int n, int Hs, int Ws, const float *dparams,
                             cudaStream_t st) {
    const TapTable &t = ctx->taps[band];
    constexpr int GP = BwdShape<G>::GP;
    const size_t smem = (size_t)(K0 * BLD + 3 * G * BLD + 16 * BLD + K0 * GP + G * GP + G * 16 + 2 * G + 16) * sizeof(float);
    LLICTI_CUDA(cudaFuncSetAttribute(cnn_backward_kernel<G, K0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!ctx->sm_count) {
        int dev = 0;
        LLICTI_CUDA(cudaGetDevice(&dev));
        LLICTI_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long tiles = (long long)((Hs * Ws + BT - 1) / BT) * n;
    // one CTA per SM (shared memory); the four sub-networks share the SMs: a quarter of them each, at least one CTA
    const int per_net = (int)std::max(1LL, std::min(tiles, (long long)ctx->sm_count / 4));
    dim3 grid(per_net, 4);
    cnn_backward_kernel<G, K0><<<grid, BNT, smem, st>>>(fplanes, Hs, Ws, t, ctx->wf32[band], dparams, n, ts->g[band]);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

template <int G>
static int launch_backward_g(llicti_ctx *ctx, TrainState *ts, int band, const float *fplanes, int n, int Hs, int Ws, const float *dparams,
                             cudaStream_t st) {
    switch (ctx->taps[band].K0) {                  // 3 x (4x4 | 3x4 + 4x3 | 4x3 + 3x4 + 4x4) taps: compile-time loop bounds
        case 48: return launch_backward_t<G, 48>(ctx, ts, band, fplanes, n, Hs, Ws, dparams, st);
        case 72: return launch_backward_t<G, 72>(ctx, ts, band, fplanes, n, Hs, Ws, dparams, st);
        case 120: return launch_backward_t<G, 120>(ctx, ts, band, fplanes, n, Hs, Ws, dparams, st);
    }
    set_error("cnn backward: unexpected layer-0 depth %d", ctx->taps[band].K0);
    return LLICTI_E_ARG;
}

template <int G, int K0>
static int launch_forward_t(llicti_ctx *ctx, int band, const float *fplanes, int n, int Hs, int Ws, float *params, cudaStream_t st) {
    const TapTable &t = ctx->taps[band];
    constexpr int GP = BwdShape<G>::GP;
    const size_t smem = (size_t)(K0 * BLD + 2 * G * BLD + K0 * GP + G * GP + G * 16) * sizeof(float);
    LLICTI_CUDA(cudaFuncSetAttribute(cnn_forward_train_kernel<G, K0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (!ctx->sm_count) {
        int dev = 0;
        LLICTI_CUDA(cudaGetDevice(&dev));
        LLICTI_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const long long tiles = (long long)((Hs * Ws + BT - 1) / BT) * n;
    const int per_net = (int)std::max(1LL, std::min(tiles, (long long)ctx->sm_count / 4));
    cnn_forward_train_kernel<G, K0><<<dim3(per_net, 4), BNT, smem, st>>>(fplanes, Hs, Ws, t, ctx->wf32[band], n, params);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

template <int G>
static int launch_forward_g(llicti_ctx *ctx, int band, const float *fplanes, int n, int Hs, int Ws, float *params, cudaStream_t st) {
    switch (ctx->taps[band].K0) {
        case 48: return launch_forward_t<G, 48>(ctx, band, fplanes, n, Hs, Ws, params, st);
        case 72: return launch_forward_t<G, 72>(ctx, band, fplanes, n, Hs, Ws, params, st);
        case 120: return launch_forward_t<G, 120>(ctx, band, fplanes, n, Hs, Ws, params, st);
    }
    set_error("cnn forward (training): unexpected layer-0 depth %d", ctx->taps[band].K0);
    return LLICTI_E_ARG;
}

// fp32 CNN of one band from the fp32 planes (forward() / training step of fp32 contexts)
int launch_cnn_forward_train(llicti_ctx *ctx, int band, const float *fplanes, int n, int Hs, int Ws, float *params, cudaStream_t st) {
    ProfScope prof_(ctx, KC_CNN, st);
    if (ctx->cfg.chs == 88) return launch_forward_g<88>(ctx, band, fplanes, n, Hs, Ws, params, st);
    if (ctx->cfg.chs == 60) return launch_forward_g<60>(ctx, band, fplanes, n, Hs, Ws, params, st);
    set_error("cnn forward (training): unsupported sub-network width %d (88 or 60)", ctx->cfg.chs);
    return LLICTI_E_ARG;
}

int launch_cnn_backward(llicti_ctx *ctx, int band, const float *fplanes, int n, int Hs, int Ws, const float *dparams, cudaStream_t st) {
    ProfScope prof_(ctx, KC_CNN, st);
    TrainState *ts;
    int rc = train_state(ctx, &ts);
    if (rc) return rc;
    if (ctx->cfg.chs == 88) return launch_backward_g<88>(ctx, ts, band, fplanes, n, Hs, Ws, dparams, st);
    if (ctx->cfg.chs == 60) return launch_backward_g<60>(ctx, ts, band, fplanes, n, Hs, Ws, dparams, st);
    set_error("cnn backward: unsupported sub-network width %d (88 or 60)", ctx->cfg.chs);
    return LLICTI_E_ARG;
}

// ---- PyTorch layouts <-> packed layouts, on the device ------------------------------------------------
// (pack_weights of api.cu, element for element; `to_packed` copies weights in, otherwise gradients out)
__global__ void l0_layout_kernel(float *__restrict__ torch_w, float *__restrict__ packed_w0, int G, int K0, int koff, int taps_br, bool to_packed) {
    const int Ch = 4 * G, GP = G + 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // over (ch, c, dy, dx) of the branch
    if (idx >= Ch * taps_br) return;
    const int ch = idx / taps_br, k = koff + idx - ch * taps_br;
    float *pk = packed_w0 + ((size_t)(ch / G) * K0 + k) * GP + ch % G;
    if (to_packed) *pk = torch_w[idx];
    else torch_w[idx] = *pk;
}

__global__ void l12_layout_kernel(float *__restrict__ torch_w1, float *__restrict__ torch_w2, float *__restrict__ packed_w1,
                                  float *__restrict__ packed_w2, int G, bool to_packed) {
    const int GP = G + 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 4 * G * G) {                                       // (g, o, i)
        const int go = idx / G, i = idx - go * G, g = go / G, o = go - g * G;
        float *pk = packed_w1 + ((size_t)g * G + i) * GP + o;
        if (to_packed) *pk = torch_w1[idx];
        else torch_w1[idx] = *pk;
    }
    if (idx < 60 * G) {                                          // (g, o, i)
        const int go = idx / G, i = idx - go * G, g = go / 15, o = go - g * 15;
        float *pk = packed_w2 + ((size_t)g * G + i) * 16 + o;
        if (to_packed) *pk = torch_w2[idx];
        else torch_w2[idx] = *pk;
    }
}

// biases: b0 = sum of the band's branch biases (every branch bias receives the gradient of b0); b1, b2 as they are
__global__ void bias_layout_kernel(float *tb0a, float *tb0b, float *tb0c, float *tb1, float *tb2, float *pb0, float *pb1, float *pb2,
                                   int G, bool to_packed) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < 4 * G) {
        if (to_packed) {
            pb0[idx] = tb0a[idx] + (tb0b ? tb0b[idx] : 0.f) + (tb0c ? tb0c[idx] : 0.f);
            pb1[idx] = tb1[idx];
        } else {
            tb0a[idx] = pb0[idx];
            if (tb0b) tb0b[idx] = pb0[idx];
            if (tb0c) tb0c[idx] = pb0[idx];
            tb1[idx] = pb1[idx];
        }
    }
    if (idx < 60) {
        if (to_packed) pb2[idx] = tb2[idx];
        else tb2[idx] = pb2[idx];
    }
}

struct BranchK { int band, kh, kw; };
static const BranchK kBr[6] = {{0, 4, 4}, {1, 3, 4}, {1, 4, 3}, {2, 4, 3}, {2, 3, 4}, {2, 4, 4}};   // order of llicti_weights.l0_*

// `tw` holds DEVICE pointers in PyTorch's layouts.  to_packed: weights into ctx->wf32; else: the gradient accumulators out.
int launch_train_layouts(llicti_ctx *ctx, const llicti_weights &tw, bool to_packed, cudaStream_t st) {
    const int G = ctx->cfg.chs;
    TrainState *ts = nullptr;
    if (!to_packed) {
        int rc = train_state(ctx, &ts);
        if (rc) return rc;
    }
    for (int band = 0; band < 3; ++band) {
        const int K0 = ctx->taps[band].K0;
        float *pw0 = to_packed ? ctx->wf32[band].w0 : ts->g[band].w0, *pb0 = to_packed ? ctx->wf32[band].b0 : ts->g[band].b0;
        float *pw1 = to_packed ? ctx->wf32[band].w1 : ts->g[band].w1, *pb1 = to_packed ? ctx->wf32[band].b1 : ts->g[band].b1;
        float *pw2 = to_packed ? ctx->wf32[band].w2 : ts->g[band].w2, *pb2 = to_packed ? ctx->wf32[band].b2 : ts->g[band].b2;
        int koff = 0;
        float *tb0[3] = {nullptr, nullptr, nullptr};
        int nb = 0;
        for (int br = 0; br < 6; ++br) {
            if (kBr[br].band != band) continue;
            LLICTI_REQUIRE(tw.l0_w[br] && tw.l0_b[br], "missing layer-0 pointer of branch %d", br);
            const int taps_br = 3 * kBr[br].kh * kBr[br].kw, total = 4 * G * taps_br;
            l0_layout_kernel<<<(total + 255) / 256, 256, 0, st>>>(const_cast<float *>(tw.l0_w[br]), pw0, G, K0, koff, taps_br, to_packed);
            koff += taps_br;
            tb0[nb++] = const_cast<float *>(tw.l0_b[br]);
            ctx->launches += 1;
        }
        LLICTI_REQUIRE(tw.l1_w[band] && tw.l1_b[band] && tw.l2_w[band] && tw.l2_b[band], "missing 1x1 pointer of band %d", band);
        l12_layout_kernel<<<(4 * G * G + 255) / 256, 256, 0, st>>>(const_cast<float *>(tw.l1_w[band]), const_cast<float *>(tw.l2_w[band]), pw1, pw2, G,
                                                                   to_packed);
        bias_layout_kernel<<<(4 * G + 255) / 256, 256, 0, st>>>(tb0[0], tb0[1], tb0[2], const_cast<float *>(tw.l1_b[band]),
                                                               const_cast<float *>(tw.l2_b[band]), pb0, pb1, pb2, G, to_packed);
        ctx->launches += 2;
    }
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
