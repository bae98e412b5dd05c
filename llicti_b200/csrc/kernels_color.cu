// Integer YCoCg-R lifting, min/max scan and the scale pyramid (forward and inverse).
// Reference: graphs/models/LLICTI_nets.py:62-88 (lifting), :137-139 (min/max), :143 (Y-127),
// :218-241 (lazyDWT with replicate padding), :446-454 / :501-509 (inverse), :571-582.
// All arithmetic is exact integer work; the kernels are HBM-bound (3 B/px in, 6 B/px out for
// the finest scale).
#include "common.cuh"

namespace llicti {

__device__ __forceinline__ void rgb_to_ycocg(int r, int g, int b, int &y, int &co, int &cg) {
    co = r - b;
    int t = b + (co >> 1);   // arithmetic shift == floor division by 2
    cg = g - t;
    y = t + (cg >> 1);
}

__device__ __forceinline__ void ycocg_to_rgb(int y, int co, int cg, int &r, int &g, int &b) {
    int t = y - (cg >> 1);
    g = cg + t;
    b = t - (co >> 1);
    r = b + co;
}

// One thread per position (r, c) of the scale-0 grid: reads the 2x2 pixel block (with the
// replicate rule for the short phases) and writes 12 int16 planes.
__global__ void __launch_bounds__(256)
split_rgb_kernel(const uint8_t *__restrict__ rgb, int H, int W, int Hs, int Ws, int H11, int W11,
                 int16_t *__restrict__ planes, int32_t *__restrict__ minmax) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    int mnCo = 1 << 20, mnCg = 1 << 20, mxCo = -(1 << 20), mxCg = -(1 << 20);
    if (c < Ws) {
        const uint8_t *base = rgb + (size_t)img * 3 * H * W;
        int16_t *out = planes + (size_t)img * 12 * Hs * Ws + (size_t)r * Ws + c;
        const size_t ps = (size_t)Hs * Ws;
        const int r1 = 2 * min(r, H11 - 1) + 1, c1 = 2 * min(c, W11 - 1) + 1;
        const int rr[4] = {2 * r, r1, 2 * r, r1};   // x00, x11, x01, x10
        const int cc[4] = {2 * c, c1, c1, 2 * c};
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const size_t o = (size_t)rr[ph] * W + cc[ph];
            const int R = base[o], G = base[o + (size_t)H * W], B = base[o + 2 * (size_t)H * W];
            int y, co, cg;
            rgb_to_ycocg(R, G, B, y, co, cg);
            out[(ph * 3 + 0) * ps] = (int16_t)(y - 127);
            out[(ph * 3 + 1) * ps] = (int16_t)co;
            out[(ph * 3 + 2) * ps] = (int16_t)cg;
            mnCo = min(mnCo, co); mxCo = max(mxCo, co);
            mnCg = min(mnCg, cg); mxCg = max(mxCg, cg);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnCo = min(mnCo, __shfl_xor_sync(0xffffffffu, mnCo, o));
        mnCg = min(mnCg, __shfl_xor_sync(0xffffffffu, mnCg, o));
        mxCo = max(mxCo, __shfl_xor_sync(0xffffffffu, mxCo, o));
        mxCg = max(mxCg, __shfl_xor_sync(0xffffffffu, mxCg, o));
    }
    if ((threadIdx.x & 31) == 0 && mnCo <= mxCo) {
        atomicMin(minmax + img * 4 + 0, mnCo);
        atomicMin(minmax + img * 4 + 1, mnCg);
        atomicMax(minmax + img * 4 + 2, mxCo);
        atomicMax(minmax + img * 4 + 3, mxCg);
    }
}

// The same for even H and W % 8 == 0 (every BASELINE shape), vectorised: a thread owns four consecutive positions of
// a scale-0 row, i.e. eight pixels of two image rows: six 8-byte loads (three channels, two rows) and twelve 8-byte
// stores (four int16 samples per plane) instead of 12 byte loads and 12 two-byte stores per position.
__global__ void __launch_bounds__(256)
split_rgb_vec_kernel(const uint8_t *__restrict__ rgb, int H, int W, int Hs, int Ws, int16_t *__restrict__ planes,
                     int32_t *__restrict__ minmax) {
    const int c4 = blockIdx.x * blockDim.x + threadIdx.x;      // group of four positions
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    int mnCo = 1 << 20, mnCg = 1 << 20, mxCo = -(1 << 20), mxCg = -(1 << 20);
    if (c4 * 4 < Ws) {
        const size_t hw = (size_t)H * W, ps = (size_t)Hs * Ws;
        const uint8_t *base = rgb + (size_t)img * 3 * hw + (size_t)(2 * r) * W + 8 * c4;
        uint2 px[2][3];                                          // [image row parity][channel]: eight pixels each
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) px[rr][ch] = __ldg(reinterpret_cast<const uint2 *>(base + (size_t)ch * hw + (size_t)rr * W));
        int16_t *out = planes + (size_t)img * 12 * ps + (size_t)r * Ws + 4 * c4;
        // phase order x00, x11, x01, x10 = (row 0, even col), (row 1, odd col), (row 0, odd col), (row 1, even col)
#pragma unroll
        for (int ph = 0; ph < 4; ++ph) {
            const int rr = (ph == 1 || ph == 3) ? 1 : 0, odd = (ph == 1 || ph == 2) ? 1 : 0;
            short y4[4], co4[4], cg4[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int byte = 2 * k + odd;                    // pixel within the eight
                const uint32_t wr = byte < 4 ? px[rr][0].x : px[rr][0].y, wg = byte < 4 ? px[rr][1].x : px[rr][1].y,
                               wb = byte < 4 ? px[rr][2].x : px[rr][2].y;
                const int sh = 8 * (byte & 3);
                int y, co, cg;
                rgb_to_ycocg((wr >> sh) & 0xFF, (wg >> sh) & 0xFF, (wb >> sh) & 0xFF, y, co, cg);
                y4[k] = (short)(y - 127); co4[k] = (short)co; cg4[k] = (short)cg;
                mnCo = min(mnCo, co); mxCo = max(mxCo, co);
                mnCg = min(mnCg, cg); mxCg = max(mxCg, cg);
            }
            auto pack = [](const short *v) {
                return make_uint2((uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16), (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16));
            };
            *reinterpret_cast<uint2 *>(out + (size_t)(ph * 3 + 0) * ps) = pack(y4);
            *reinterpret_cast<uint2 *>(out + (size_t)(ph * 3 + 1) * ps) = pack(co4);
            *reinterpret_cast<uint2 *>(out + (size_t)(ph * 3 + 2) * ps) = pack(cg4);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mnCo = min(mnCo, __shfl_xor_sync(0xffffffffu, mnCo, o));
        mnCg = min(mnCg, __shfl_xor_sync(0xffffffffu, mnCg, o));
        mxCo = max(mxCo, __shfl_xor_sync(0xffffffffu, mxCo, o));
        mxCg = max(mxCg, __shfl_xor_sync(0xffffffffu, mxCg, o));
    }
    if ((threadIdx.x & 31) == 0 && mnCo <= mxCo) {
        atomicMin(minmax + img * 4 + 0, mnCo);
        atomicMin(minmax + img * 4 + 1, mnCg);
        atomicMax(minmax + img * 4 + 2, mxCo);
        atomicMax(minmax + img * 4 + 3, mxCg);
    }
}

// Scale s+1 from the x00 planes of scale s (int16 for the coder, fp32 for the rate-estimation path).
template <typename T>
__global__ void __launch_bounds__(256)
split_x00_kernel(const T *__restrict__ prev, int Hp, int Wp, int Hs, int Ws, int H11, int W11,
                 T *__restrict__ planes) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    if (c >= Ws) return;
    const T *base = prev + (size_t)img * 12 * Hp * Wp;   // channels 0..2 = x00 of the finer scale
    T *out = planes + (size_t)img * 12 * Hs * Ws + (size_t)r * Ws + c;
    const size_t ps = (size_t)Hs * Ws, pp = (size_t)Hp * Wp;
    const int r1 = 2 * min(r, H11 - 1) + 1, c1 = 2 * min(c, W11 - 1) + 1;
    const int rr[4] = {2 * r, r1, 2 * r, r1};
    const int cc[4] = {2 * c, c1, c1, 2 * c};
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const size_t o = (size_t)rr[ph] * Wp + cc[ph];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) out[(ph * 3 + ch) * ps] = base[o + ch * pp];
    }
}

__global__ void init_minmax_kernel(int32_t *minmax, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        minmax[i * 4 + 0] = 1 << 20;
        minmax[i * 4 + 1] = 1 << 20;
        minmax[i * 4 + 2] = -(1 << 20);
        minmax[i * 4 + 3] = -(1 << 20);
    }
}

__global__ void minmax16_kernel(const int32_t *__restrict__ mm, int16_t *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {   // header words [minY, minCo, minCg, maxY, maxCo, maxCg], LLICTI_nets.py:137-139
        out[i * 6 + 0] = 0;
        out[i * 6 + 1] = (int16_t)mm[i * 4 + 0];
        out[i * 6 + 2] = (int16_t)mm[i * 4 + 1];
        out[i * 6 + 3] = 255;
        out[i * 6 + 4] = (int16_t)mm[i * 4 + 2];
        out[i * 6 + 5] = (int16_t)mm[i * 4 + 3];
    }
}

__global__ void minmax32_kernel(const int16_t *__restrict__ mm16, int32_t *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        out[i * 4 + 0] = mm16[i * 6 + 1];
        out[i * 4 + 1] = mm16[i * 6 + 2];
        out[i * 4 + 2] = mm16[i * 6 + 4];
        out[i * 4 + 3] = mm16[i * 6 + 5];
    }
}

// Coarsest x00 from the header's raw RGB (LLICTI_nets.py:429-430, :444).
__global__ void x00_header_kernel(const uint8_t *__restrict__ x00, int h, int w, int16_t *__restrict__ planes, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (i >= h * w) return;
    const uint8_t *base = x00 + (size_t)img * 3 * h * w;
    int y, co, cg;
    rgb_to_ycocg(base[i], base[i + h * w], base[i + 2 * h * w], y, co, cg);
    int16_t *out = planes + (size_t)img * 12 * h * w + i;
    out[0] = (int16_t)(y - 127);
    out[(size_t)h * w] = (int16_t)co;
    out[2 * (size_t)h * w] = (int16_t)cg;
}

// Inverse lazy DWT of scale s into the x00 planes of scale s-1 (cropped to Ht x Wt).
__global__ void __launch_bounds__(256)
interleave_kernel(const int16_t *__restrict__ from, int Hs, int Ws, int16_t *__restrict__ to, int Ht, int Wt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    if (c >= Wt) return;
    const int ph = (r & 1) ? ((c & 1) ? 1 : 3) : ((c & 1) ? 2 : 0);
    const size_t ps = (size_t)Hs * Ws, pt = (size_t)Ht * Wt;
    const int16_t *src = from + (size_t)img * 12 * ps + (size_t)(ph * 3) * ps + (size_t)(r >> 1) * Ws + (c >> 1);
    int16_t *dst = to + (size_t)img * 12 * pt + (size_t)r * Wt + c;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) dst[ch * pt] = src[ch * ps];
}

// Final inverse lazy DWT + Y+127 + inverse lifting -> planar RGB uint8.
__global__ void __launch_bounds__(256)
merge_rgb_kernel(const int16_t *__restrict__ from, int Hs, int Ws, uint8_t *__restrict__ rgb, int H, int W) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    if (c >= W) return;
    const int ph = (r & 1) ? ((c & 1) ? 1 : 3) : ((c & 1) ? 2 : 0);
    const size_t ps = (size_t)Hs * Ws;
    const int16_t *src = from + (size_t)img * 12 * ps + (size_t)(ph * 3) * ps + (size_t)(r >> 1) * Ws + (c >> 1);
    int R, G, B;
    ycocg_to_rgb(src[0] + 127, src[ps], src[2 * ps], R, G, B);
    uint8_t *dst = rgb + (size_t)img * 3 * H * W + (size_t)r * W + c;
    dst[0] = (uint8_t)R;
    dst[(size_t)H * W] = (uint8_t)G;
    dst[2 * (size_t)H * W] = (uint8_t)B;
}

int launch_color_split(llicti_ctx *ctx, const Plan &p, const uint8_t *rgb, int n, int16_t *const *planes,
                       int32_t *minmax, cudaStream_t st) {
    ProfScope prof_(ctx, KC_SPLIT, st);
    const llicti_geom &g = p.g;
    init_minmax_kernel<<<(n + 127) / 128, 128, 0, st>>>(minmax, n);
    const bool aligned = (reinterpret_cast<uintptr_t>(rgb) % 8 == 0) && (reinterpret_cast<uintptr_t>(planes[0]) % 8 == 0);
    if (g.H % 2 == 0 && g.W % 8 == 0 && aligned) {
        const int groups = g.Ws[0] / 4, threads = groups >= 256 ? 256 : 64;
        dim3 grid((groups + threads - 1) / threads, g.Hs[0], n);
        split_rgb_vec_kernel<<<grid, threads, 0, st>>>(rgb, g.H, g.W, g.Hs[0], g.Ws[0], planes[0], minmax);
    } else {
        dim3 grid((g.Ws[0] + 255) / 256, g.Hs[0], n);
        split_rgb_kernel<<<grid, 256, 0, st>>>(rgb, g.H, g.W, g.Hs[0], g.Ws[0], g.H / 2, g.W / 2, planes[0], minmax);
    }
    ctx->launches += 2;
    for (int s = 1; s < g.num_scales; ++s) {
        dim3 grid((g.Ws[s] + 255) / 256, g.Hs[s], n);
        split_x00_kernel<int16_t><<<grid, 256, 0, st>>>(planes[s - 1], g.Hs[s - 1], g.Ws[s - 1], g.Hs[s], g.Ws[s],
                                                g.Hs[s - 1] / 2, g.Ws[s - 1] / 2, planes[s]);
        ctx->launches += 1;
    }
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

// ---- rate-estimation path (LLICTI.forward): the FLOAT lifting of LLICTI_nets.py:40-49 --------------------
// x = k / 255 (ToTensor: IEEE division on the host), Co = R - B, t = B + round(Co * 255 / 2) / 255,
// Cg = G - t, Y = t + round(Cg * 255 / 2) / 255, Y -= 127/255 (:110), every step in fp32 with round-half-even --
// operation for operation what the reference's tensor expressions evaluate.  Not the integer transform of the
// coder (that one floors): odd differences land on other integers, and a product like 1.4999999 rounds down.
// One thread per scale-0 position; writes the fp32 planes the likelihood reads and the int16 planes
// (value * 255, an integer up to fp32 noise) the CNN kernels read.  H and W are multiples of 2^S: no padding.
__global__ void __launch_bounds__(256)
split_rgb_float_kernel(const uint8_t *__restrict__ rgb, int H, int W, int Hs, int Ws, NumericsProfile np,
                       float *__restrict__ fplanes, int16_t *__restrict__ planes) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    if (c >= Ws) return;
    const uint8_t *base = rgb + (size_t)img * 3 * H * W;
    const size_t ps = (size_t)Hs * Ws;
    float *fo = fplanes + (size_t)img * 12 * ps + (size_t)r * Ws + c;
    int16_t *io = planes + (size_t)img * 12 * ps + (size_t)r * Ws + c;
    const int rr[4] = {2 * r, 2 * r + 1, 2 * r, 2 * r + 1};   // x00, x11, x01, x10
    const int cc[4] = {2 * c, 2 * c + 1, 2 * c + 1, 2 * c};
    const float mean_y = (float)(127.0 / 255.0);
    auto div255 = [&](float v) { return np.div255_recip ? __fmul_rn(v, 1.0f / 255.0f) : __fdiv_rn(v, 255.0f); };
#pragma unroll
    for (int ph = 0; ph < 4; ++ph) {
        const size_t o = (size_t)rr[ph] * W + cc[ph];
        const float R = __fdiv_rn((float)base[o], 255.0f), G = __fdiv_rn((float)base[o + (size_t)H * W], 255.0f),
                    B = __fdiv_rn((float)base[o + 2 * (size_t)H * W], 255.0f);
        const float co = __fsub_rn(R, B);
        const float t = __fadd_rn(B, div255(rintf(__fmul_rn(__fmul_rn(co, 255.0f), 0.5f))));
        const float cg = __fsub_rn(G, t);
        const float y = __fsub_rn(__fadd_rn(t, div255(rintf(__fmul_rn(__fmul_rn(cg, 255.0f), 0.5f)))), mean_y);
        const float v[3] = {y, co, cg};
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            fo[(ph * 3 + ch) * ps] = v[ch];
            io[(ph * 3 + ch) * ps] = (int16_t)__float2int_rn(__fmul_rn(v[ch], 255.0f));
        }
    }
}

int launch_color_split_float(llicti_ctx *ctx, const Plan &p, const uint8_t *rgb, int n, float *const *fplanes,
                             int16_t *const *planes, cudaStream_t st) {
    ProfScope prof_(ctx, KC_SPLIT, st);
    const llicti_geom &g = p.g;
    {
        dim3 grid((g.Ws[0] + 255) / 256, g.Hs[0], n);
        split_rgb_float_kernel<<<grid, 256, 0, st>>>(rgb, g.H, g.W, g.Hs[0], g.Ws[0], ctx->num, fplanes[0], planes[0]);
    }
    ctx->launches += 1;
    for (int s = 1; s < g.num_scales; ++s) {
        dim3 grid((g.Ws[s] + 255) / 256, g.Hs[s], n);
        split_x00_kernel<float><<<grid, 256, 0, st>>>(fplanes[s - 1], g.Hs[s - 1], g.Ws[s - 1], g.Hs[s], g.Ws[s],
                                                       g.Hs[s - 1] / 2, g.Ws[s - 1] / 2, fplanes[s]);
        split_x00_kernel<int16_t><<<grid, 256, 0, st>>>(planes[s - 1], g.Hs[s - 1], g.Ws[s - 1], g.Hs[s], g.Ws[s],
                                                         g.Hs[s - 1] / 2, g.Ws[s - 1] / 2, planes[s]);
        ctx->launches += 2;
    }
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

// Vectorised form for even H and W % 8 == 0: a thread owns eight consecutive pixels of an image row -- four samples of
// each of the two phases that alternate along the row, three channels: six 8-byte loads, three 8-byte stores.
__global__ void __launch_bounds__(256)
merge_rgb_vec_kernel(const int16_t *__restrict__ from, int Hs, int Ws, uint8_t *__restrict__ rgb, int H, int W) {
    const int c8 = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    const int img = blockIdx.z;
    if (c8 * 8 >= W) return;
    const size_t ps = (size_t)Hs * Ws, hw = (size_t)H * W;
    const int ph_even = (r & 1) ? 3 : 0, ph_odd = (r & 1) ? 1 : 2;          // phase of the even / odd columns of this row
    const int16_t *src = from + (size_t)img * 12 * ps + (size_t)(r >> 1) * Ws + 4 * c8;
    uint2 ev[3], od[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        ev[ch] = __ldg(reinterpret_cast<const uint2 *>(src + (size_t)(ph_even * 3 + ch) * ps));
        od[ch] = __ldg(reinterpret_cast<const uint2 *>(src + (size_t)(ph_odd * 3 + ch) * ps));
    }
    uint32_t outw[3][2] = {};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int j = k >> 1;                                               // sample within the four
        const uint2 *v = (k & 1) ? od : ev;
        auto pick = [&](const uint2 &q) { const uint32_t w = j < 2 ? q.x : q.y; return (int)(short)((j & 1) ? (w >> 16) : (w & 0xFFFF)); };
        int R, G, B;
        ycocg_to_rgb(pick(v[0]) + 127, pick(v[1]), pick(v[2]), R, G, B);
        outw[0][k >> 2] |= (uint32_t)(R & 0xFF) << (8 * (k & 3));
        outw[1][k >> 2] |= (uint32_t)(G & 0xFF) << (8 * (k & 3));
        outw[2][k >> 2] |= (uint32_t)(B & 0xFF) << (8 * (k & 3));
    }
    uint8_t *dst = rgb + (size_t)img * 3 * hw + (size_t)r * W + 8 * c8;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) *reinterpret_cast<uint2 *>(dst + (size_t)ch * hw) = make_uint2(outw[ch][0], outw[ch][1]);
}

int launch_merge_color(llicti_ctx *ctx, const Plan &p, const int16_t *planes0, int n, uint8_t *rgb, cudaStream_t st) {
    ProfScope prof_(ctx, KC_MERGE, st);
    const llicti_geom &g = p.g;
    const bool aligned = (reinterpret_cast<uintptr_t>(rgb) % 8 == 0) && (reinterpret_cast<uintptr_t>(planes0) % 8 == 0);
    if (g.H % 2 == 0 && g.W % 8 == 0 && aligned) {
        const int groups = g.W / 8, threads = groups >= 256 ? 256 : 64;
        dim3 grid((groups + threads - 1) / threads, g.H, n);
        merge_rgb_vec_kernel<<<grid, threads, 0, st>>>(planes0, g.Hs[0], g.Ws[0], rgb, g.H, g.W);
        ctx->launches += 1;
        LLICTI_CUDA(cudaGetLastError());
        return LLICTI_OK;
    }
    dim3 grid((g.W + 255) / 256, g.H, n);
    merge_rgb_kernel<<<grid, 256, 0, st>>>(planes0, g.Hs[0], g.Ws[0], rgb, g.H, g.W);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_x00_from_header(llicti_ctx *ctx, const Plan &p, const uint8_t *x00_rgb, int n, int16_t *planes_last,
                           cudaStream_t st) {
    ProfScope prof_(ctx, KC_MERGE, st);
    const llicti_geom &g = p.g;
    const int s = g.num_scales - 1;
    dim3 grid((g.Hs[s] * g.Ws[s] + 127) / 128, n);
    x00_header_kernel<<<grid, 128, 0, st>>>(x00_rgb, g.Hs[s], g.Ws[s], planes_last, n);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_interleave(llicti_ctx *ctx, const Plan &p, int scale_from, const int16_t *planes_from, int16_t *planes_to,
                      int n, cudaStream_t st) {
    ProfScope prof_(ctx, KC_MERGE, st);
    const llicti_geom &g = p.g;
    const int s = scale_from, t = scale_from - 1;
    dim3 grid((g.Ws[t] + 255) / 256, g.Hs[t], n);
    interleave_kernel<<<grid, 256, 0, st>>>(planes_from, g.Hs[s], g.Ws[s], planes_to, g.Hs[t], g.Ws[t]);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_minmax16(llicti_ctx *ctx, const int32_t *minmax, int16_t *minmax16, int n, cudaStream_t st) {
    ProfScope prof_(ctx, KC_MERGE, st);
    minmax16_kernel<<<(n + 127) / 128, 128, 0, st>>>(minmax, minmax16, n);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_minmax32(llicti_ctx *ctx, const int16_t *minmax16, int32_t *minmax, int n, cudaStream_t st) {
    ProfScope prof_(ctx, KC_MERGE, st);
    minmax32_kernel<<<(n + 127) / 128, 128, 0, st>>>(minmax16, minmax, n);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
