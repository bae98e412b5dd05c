// Warp-cooperative evaluation of the integer GMM CDF: the 32 lanes of a warp evaluate 32 table
// entries of ONE position at once (decode-side slow path and the legacy one-warp-per-chain decoder).
#pragma once

#include "common.cuh"
#include "gmm.cuh"

namespace llicti {

constexpr unsigned kFull = 0xffffffffu;

// 16-bit cumulative count from the coder state; floor((value-low+1)*2^16 - 1) / span) in fp64
// is exact here: num < 2^49, span <= 2^32, and a non-integer quotient is at least 2^-32 away
// from an integer while the fp64 quotient is within 2^-36 of the true one.
__device__ __forceinline__ uint32_t target_fp64(uint32_t low, uint32_t high, uint32_t value) {
    const double span = (double)(high - low) + 1.0;
    const double num = ((double)(value - low) + 1.0) * 65536.0 - 1.0;
    return __double2uint_rd(num / span) & 0xFFFFu;
}

struct WarpParams {   // the 60 network outputs of one position, two per lane
    float p0, p1;
    __device__ __forceinline__ void load(const float *__restrict__ pp, size_t P, size_t pidx, int lane) {
        p0 = pp[(size_t)lane * P + pidx];
        p1 = lane < kParamCh - 32 ? pp[(size_t)(lane + 32) * P + pidx] : 0.f;
    }
    __device__ __forceinline__ float get(int ch) const {
        return ch < 32 ? __shfl_sync(kFull, p0, ch) : __shfl_sync(kFull, p1, ch - 32);
    }
};

__device__ __forceinline__ void warp_channel(const WarpParams &wp, int clr, int y0, int y1, const NumericsProfile &np,
                                             GmmChannel &c) {
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        c.sigma[m] = wp.get(clr * kM + m);
        c.mu[m] = wp.get((3 + clr) * kM + m);
        c.w[m] = wp.get((6 + clr) * kM + m);
    }
    if (clr == 1) {
        const float f0 = div255((float)y0, np);
#pragma unroll
        for (int m = 0; m < kM; ++m) c.mu[m] = __fadd_rn(c.mu[m], __fmul_rn(wp.get(9 * kM + m), f0));
    } else if (clr == 2) {
        const float f0 = div255((float)y0, np), f1 = div255((float)y1, np);
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            const float u = __fadd_rn(__fmul_rn(wp.get(10 * kM + m), f0), __fmul_rn(wp.get(11 * kM + m), f1));
            c.mu[m] = __fadd_rn(c.mu[m], u);
        }
    }
    gmm_prepare(c, np);
}

// Largest m in [0, Lp-2] with q(m) <= target, found by rounds of 32 parallel probes: first a
// unit-stride window around the predicted value, then (rarely) a coarse round over what is left
// of [0, Lp-1] and a final unit-stride round.  Entry Lp-1 acts as the 0x10000 sentinel.
// `first_base` >= 0 overrides the first window (the decode slow path knows on which side of the
// predicted window the symbol lies and starts right next to it).
__device__ __forceinline__ int warp_search(const GmmChannel &c, const CdfGrid &g, uint32_t target,
                                           const NumericsProfile &np, int lane, uint32_t &c_low, uint32_t &c_high,
                                           int first_base = -1) {
    const int last = g.Lp - 1;
    int lo = 0, hi = last;
    int base;
    if (first_base >= 0) {
        base = min(first_base, max(last - 31, 0));
    } else {
        float mean = 0.f;
#pragma unroll
        for (int m = 0; m < kM; ++m) mean = fmaf(c.w[m], c.mu[m], mean);
        const int kc = __float2int_rn(mean * 255.0f) - g.min_val;
        base = min(max(kc - 15, 0), max(last - 31, 0));
    }
    int stride = 1;
    for (;;) {
        const int k = base + lane * stride;
        const uint32_t qv = k < last ? cdf_q(c, g, k, np) : 0x10000u;
        const unsigned le = __ballot_sync(kFull, qv <= target);
        const int cnt = __popc(le);
        if (cnt == 0) {
            if (base == 0 && stride == 1) {      // target below q(0): torchac's search returns 0
                c_low = __shfl_sync(kFull, qv, 0);
                c_high = __shfl_sync(kFull, qv, 1);
                return 0;
            }
            hi = base == 0 ? stride : base;
        } else if (cnt == 32) {
            lo = base + 31 * stride;
        } else {
            lo = base + (cnt - 1) * stride;
            hi = min(base + cnt * stride, last);
            if (stride == 1) {
                c_low = __shfl_sync(kFull, qv, cnt - 1);
                c_high = __shfl_sync(kFull, qv, cnt);
                return lo;
            }
        }
        stride = max((hi - lo + 30) / 31, 1);
        base = lo;
    }
}

}  // namespace llicti
