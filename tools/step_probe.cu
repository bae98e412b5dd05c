// Serial-step latency of the decoder: one warp runs the product's consume_chain() over synthetic window
// items (every window covers the whole 31-symbol alphabet, so no slow path) and reports cycles per symbol.
// Built once per variant:  nvcc -DLLICTI_CLZ_I2F=c ... -o tools/_bin/step_probe_c tools/step_probe.cu
#include "../llicti_b200/csrc/kernels_decode.cu"

namespace llicti {
void set_error(const char *, ...) {}
int launch_cnn_tc(llicti_ctx *, int, const int16_t *, int, int, int, float *, cudaStream_t, int, int) { return 0; }

__global__ void __launch_bounds__(128) step_probe_kernel(const uint4 *items, const uint8_t *stream, uint32_t stream_len, int n_sym,
                                                         int16_t *out, long long *cyc, int warps) {
    __shared__ __align__(16) int li_buf[4][kLiBuf];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w >= warps) return;
    const CdfGrid g = make_grid(0, 30);     // Lp = 32: symbols 0..30
    NumericsProfile np; np.div255_recip = 1; np.sum_ilp4 = 1;
    const ChainCtx cx = {nullptr, out, (size_t)n_sym, 1, 1 << 30, 1 << 30, 0, 0, 0, 1};
    const long long t0 = clock64();
    consume_chain<false>(cx, n_sym, g, 0, np, out + (size_t)w * n_sym, items, nullptr, stream + 64 * w, stream_len - 64 * w, lane, li_buf[w]);
    const long long t1 = clock64();
    if (lane == 0) cyc[w] = t1 - t0;
}
}  // namespace llicti

int main() {
    using namespace llicti;
    const int n_sym = 1 << 16, n_items = n_sym / 32;
    const size_t item_u16 = (size_t)kItemU4 * 8;
    std::vector<uint16_t> h((size_t)n_items * item_u16, 0);
    // item layout: uint4 [4][32 lanes]; lane l, element e of its v-th uint4 = entry l of step 8 v + e (entry 31 =
    // the alphabet's end = 2^16, stored as 0); then the 32 window bases (0 here)
    for (int it = 0; it < n_items; ++it)
        for (int v = 0; v < 4; ++v)
            for (int l = 0; l < 32; ++l)
                for (int e = 0; e < 8; ++e)
                    h[(size_t)it * item_u16 + ((size_t)v * 32 + l) * 8 + e] = l < 31 ? (uint16_t)(l * 2114) : (uint16_t)0;
    std::vector<uint8_t> hs(1 << 20);
    uint64_t s = 88172645463325252ull;
    for (auto &b : hs) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; b = (uint8_t)(s >> 24); }
    uint4 *items; uint8_t *stream; int16_t *out; long long *cyc;
    cudaMalloc(&items, h.size() * 2); cudaMalloc(&stream, hs.size()); cudaMalloc(&out, (size_t)4 * n_sym * 2); cudaMalloc(&cyc, 64);
    cudaMemcpy(items, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(stream, hs.data(), hs.size(), cudaMemcpyHostToDevice);
    for (int warps = 1; warps <= 4; warps += 3) {
        long long hc[4] = {0, 0, 0, 0}, best = 1ll << 60;
        for (int r = 0; r < 3; ++r) {
            step_probe_kernel<<<1, 128>>>(items, stream, (uint32_t)hs.size() / 2, n_sym, out, cyc, warps);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
            if (hc[0] < best) best = hc[0];
        }
        unsigned long long st[8];
        cudaMemcpyFromSymbol(st, g_decode_stats, sizeof(st));
        std::vector<int16_t> ho(n_sym);
        cudaMemcpy(ho.data(), out, n_sym * 2, cudaMemcpyDeviceToHost);
        unsigned long long cs = 0;
        for (int i = 0; i < n_sym; ++i) cs = cs * 1000003ull + (uint16_t)ho[i];
        printf("clz_i2f %d warps %d: %.1f cycles per symbol (slow %llu, redone %llu, checksum %016llx)\n",
               LLICTI_CLZ_I2F, warps, (double)best / n_sym, st[0], st[3], cs);
    }
    return 0;
}
