// Rate estimation without coding: -log2 of the probability mass the GMM gives every sample
// (LLICTIEntropyModel4.get_self_infos, graphs/models/LLICTI_nets.py:827-935, clr_joint_mode 2 branch :855-866;
// GaussianConditionalLosslessGMM.forward, graphs/layers/entropy_layer_nets.py:160-183 with _likelihood_fk
// :121-139).  It is what LLICTI.forward returns per scale ([B, 9, Hs, Ws]: band-major, Y Co Cg) and what the
// reference's validate() sums into bits per pixel.
#include "common.cuh"
#include "gmm.cuh"

namespace llicti {

// One thread per position of a band: the three colour channels in turn (the means of Co and Cg are coupled to the
// float samples of Y and Co of the same position, :858-860).
__global__ void __launch_bounds__(128)
self_info_kernel(const float *__restrict__ params, const float *__restrict__ fplanes, int band, int P, NumericsProfile np,
                 float *__restrict__ sinfo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (i >= P) return;
    const float *pp = params + (size_t)img * kParamCh * P + i;
    const float *yt = fplanes + (size_t)img * 12 * P + (size_t)(3 * (band + 1)) * P + i;
    float *out = sinfo + (size_t)img * 9 * P + (size_t)(3 * band) * P + i;
    const float y0 = yt[0], y1 = yt[P], y2 = yt[2 * (size_t)P];
    const float half = (float)(0.5 / 255.0), sb = (float)(0.11 / 255.0);
    const float v[3] = {y0, y1, y2};
#pragma unroll
    for (int clr = 0; clr < 3; ++clr) {
        float sg[kM], mu[kM], w[kM], t[kM];
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            sg[m] = fmaxf(pp[(size_t)(clr * kM + m) * P], sb);
            mu[m] = pp[(size_t)((3 + clr) * kM + m) * P];
            w[m] = fmaxf(pp[(size_t)((6 + clr) * kM + m) * P], 1e-6f);
            if (clr == 1) mu[m] = __fadd_rn(mu[m], __fmul_rn(pp[(size_t)(9 * kM + m) * P], y0));
            if (clr == 2)
                mu[m] = __fadd_rn(mu[m], __fadd_rn(__fmul_rn(pp[(size_t)(10 * kM + m) * P], y0), __fmul_rn(pp[(size_t)(11 * kM + m) * P], y1)));
        }
        const float den = sum5(w, np);                                   // weights / torch.sum(weights) (:177; no epsilon here)
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            const float a = fabsf(__fsub_rn(v[clr], mu[m]));
            const float up = __fmul_rn(0.5f, erfcf(__fmul_rn(-0.70710678118654752440f, __fdiv_rn(__fsub_rn(half, a), sg[m]))));
            const float lo = __fmul_rn(0.5f, erfcf(__fmul_rn(-0.70710678118654752440f, __fdiv_rn(__fsub_rn(-half, a), sg[m]))));
            t[m] = __fmul_rn(__fdiv_rn(w[m], den), __fsub_rn(up, lo));
        }
        const float lik = fmaxf(sum5(t, np), 1e-9f);                     // likelihood_lower_bound
        out[(size_t)clr * P] = -log2f(lik);
    }
}

int launch_self_info(llicti_ctx *ctx, const float *params, const float *fplanes, int band, int n, int P, float *sinfo,
                     cudaStream_t st) {
    ProfScope prof_(ctx, KC_BOUNDS, st);
    dim3 grid((P + 127) / 128, n);
    self_info_kernel<<<grid, 128, 0, st>>>(params, fplanes, band, P, ctx->num, sinfo);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
