// Decode side of the entropy layer: mean coupling -> integer GMM CDF -> arithmetic decoding
// (LLICTIEntropyLayer.decompress inner loop, graphs/models/LLICTI_nets.py:463-498, with
// torchac.decode_int16_normalized_cdf at :492-493).
//
// The reference builds the dense H x W x Lp table for every stream and lets one CPU thread
// binary-search it symbol by symbol.  Here the work is split by what depends on the coder state:
//
//   producers  (any number of warps, no coder state): for every symbol, the 32 exact table
//              entries q(base .. base+31) around the predicted value -- a "window" of 31 complete
//              symbols, one 64-byte row.  32 symbols of one chain form an item: 2 KB of rows, the 32
//              window bases and (piped schedules) the prepared GMM channel of every step.
//   consumers  (one warp per coded chain, the serial part): per symbol two high-word multiplies
//              and a compare per lane, a ballot, four shuffles and the interval update.  No erfc, no
//              division, no table search on the critical path.  A symbol outside its window
//              (rare) takes the slow path: the analytic search of `warp_search`.
//
// Three schedules use the same two device functions:
//   split      six launches per band (window Y, consume Y, window Co, ...): any number of
//              substreams; the Y -> Co -> Cg coupling is ordered by the launches; two half batches
//              run one kernel out of phase on two streams.
//   piped      one launch per band when every stream is a single torchac chain (sub_len = 0):
//              three consumer warps per image (Y, Co, Cg) run concurrently with the producers in
//              one grid of one-warp CTAs that are all co-resident.  Consumers publish decoded
//              samples by writing them over a sentinel (data = flag, no fence on the serial
//              path); producers publish items through per-item flags.  The serial chain per band
//              is n_sym steps instead of 3 * n_sym.
//   wavefront  (default for sub_len = 0) the three bands of a scale run concurrently a strip of
//              rows apart: a consumer and a producer kernel side by side per time step.
#include "common.cuh"
#include "gmm.cuh"
#include "rangecoder.cuh"
#include "warp_gmm.cuh"

#include <algorithm>
#include <cstdlib>

namespace llicti {

#ifndef LLICTI_WAIT_SLEEP
#define LLICTI_WAIT_SLEEP 0      // ns a consumer sleeps between two looks at an item flag (it has its scheduler to itself)
#endif

__device__ unsigned long long g_decode_stats[8];   // 0 slow-path symbols, 1 consumer flag polls, 2 longest consumer run (cycles), 3 chunks redone carefully,
                                                   // 4 cycles consumers waited for items, 5 consumer cycles, 6 consumer runs, 7 cycles in redone chunks (piped schedules)

__device__ unsigned long long g_wave_dbg[8];   // LLICTI_WAVE_DEBUG: per-launch consumer timing
__global__ void wave_dbg_kernel(int T, int chains) {
    printf("[wave] T=%d chains=%d max=%.3f ms avg=%.3f ms wait=%.3f (Y %.3f Co %.3f Cg %.3f per chain of the channel, first item %.3f) redo=%.3f\n", T, chains,
           g_wave_dbg[0] / 1.965e6, g_wave_dbg[1] / 1.965e6 / chains, g_wave_dbg[2] / 1.965e6 / chains, g_wave_dbg[4] / 1.965e6 / (chains / 3),
           g_wave_dbg[5] / 1.965e6 / (chains / 3), g_wave_dbg[6] / 1.965e6 / (chains / 3), g_wave_dbg[7] / 1.965e6 / chains, g_wave_dbg[3] / 1.965e6 / chains);
    for (int i = 0; i < 8; ++i) g_wave_dbg[i] = 0;
}

constexpr int kStageU16 = 32 * 34 + 32 * 6 * 8;   // per warp: window staging (17-word pitch) + prepared channels of 32 steps
constexpr int kWin = 31;                 // complete symbols per window (32 table entries: lane l holds q(base + l))
constexpr int kLiBuf = 32;               // per consumer warp: the window-relative symbols of the item in flight
constexpr int kItemU4 = 132 + 192;       // uint4 per item: [4][32 lanes] window rows, 32 x u16 window bases, then (piped schedules
                                         // only) the prepared GMM channel of every step, [32][6] float4, for the consumer's slow path
constexpr int16_t kSentinel = (int16_t)0x8080;   // memset(0x80): never a sample value (|v| <= 255)


struct DecodeGeom {
    int Hs, Ws, crop_h, crop_w, band, padH, padW;
    int n_sym;            // crop_h * crop_w
    int S;                // substreams (chains) per stream of this band
    int max_steps;        // ceil(n_sym / S)
    int items_per_chain;  // ceil(max_steps / 32)
    int sub_first[3];     // first substream of the Y / Co / Cg stream among the image's substreams
};

// ---- strong (L2-coherent) accesses for data exchanged between CTAs of one running grid --------
__device__ __forceinline__ int ld_relaxed_s16(const int16_t *p) {
    short v;
    asm volatile("ld.relaxed.gpu.global.s16 %0, [%1];" : "=h"(v) : "l"(p) : "memory");
    return (int)v;
}
__device__ __forceinline__ void st_relaxed_s16(int16_t *p, int v) {
    asm volatile("st.relaxed.gpu.global.b16 [%0], %1;" ::"l"(p), "h"((short)v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Decoded symbol of coding-order index i of one (image, channel); in the piped schedule the
// array starts as sentinels and the consumer's store is the publication (data = flag).
// Waits inside the piped grid are bounded: a wait that outlives any legitimate kernel run time (a
// programming error, or a grid that is not co-resident) traps instead of hanging the GPU.
// The kernels of the piped and wavefront schedules wait for one another (consumers for items, producers for
// decoded symbols), which only makes progress while both sides are resident.  CUDA does not promise that -- another
// process on the GPU, a tool that serialises launches, a smaller part -- so no wait is unbounded and none traps: a
// wait that outlives g_max_polls raises g_abort, every waiter sees the flag and leaves, the launch sequence ends
// with abort_to_status_kernel (status = LLICTI_E_TIMEOUT), and the host re-runs the decode with the split
// schedule, whose kernels never wait on each other (api.cu: llicti_decode_host).
__device__ uint32_t g_max_polls = 1u << 25;    // x >= 100 ns per poll: several seconds (LLICTI_TEST_POLL_LIMIT shortens it in tests)
__device__ int g_abort = 0;
__device__ int g_test_starve = 0;              // LLICTI_TEST_STARVE: producers leave at once (a producer kernel that never runs)

__device__ __forceinline__ bool aborted() { return *(volatile int *)&g_abort != 0; }
__device__ __forceinline__ bool give_up(uint32_t polls) {
    if ((polls & 1023u) == 1023u && aborted()) return true;
    if (polls > *(volatile uint32_t *)&g_max_polls) { atomicExch(&g_abort, 1); return true; }
    return false;
}

__global__ void abort_to_status_kernel(int32_t *status) {
    if (g_abort) { atomicExch(status, LLICTI_E_TIMEOUT); g_abort = 0; }
}

template <bool kPipe>
__device__ __forceinline__ int read_symbol(const int16_t *p) {
    if (!kPipe) return (int)*p;
    int v = ld_relaxed_s16(p);
    for (uint32_t polls = 0; v == (int)kSentinel; ++polls) {       // not decoded yet
        if (give_up(polls)) return 0;                             // the launch is being abandoned: any value will do
        __nanosleep(200);
        v = ld_relaxed_s16(p);
    }
    return v;
}

// false: the wait was abandoned (g_abort)
__device__ __forceinline__ bool wait_flag(const uint32_t *flag, unsigned long long &polls, long long &waited) {
    if (ld_relaxed_u32(flag) != 0u) return true;
    const long long t0 = clock64();
    for (uint32_t spins = 0; ld_relaxed_u32(flag) == 0u; ++spins) {
        if (give_up(spins >> 2)) return false;
        ++polls;
        if (LLICTI_WAIT_SLEEP) __nanosleep(LLICTI_WAIT_SLEEP);
    }
    waited += clock64() - t0;
    return true;
}

// ------------------------------------------------------------------------------------------
// Producer: the windows of 32 consecutive steps of one chain.  Item layout: uint4 [4][32 lanes],
// lane l holds its own entries q(base_s + l) of steps s = 8 v + e in element e of its v-th uint4;
// the entry of the alphabet's end, q(Lp - 1) = 2^16, and everything past it is stored as 0.  The
// 32 bases follow as u16.  A window holds the 31 symbols base .. base+30 (fewer for alphabets
// below 31 symbols, where base = 0).
// `syms` = compact symbol arrays of this image's band, [3][sym_cap] in coding order.
// ------------------------------------------------------------------------------------------
template <bool kPipe>
__device__ __forceinline__ void produce_item(const float *__restrict__ pp, const int16_t *syms, size_t sym_cap, size_t P,
                                             const DecodeGeom &dg, int clr, const int (&lo)[3], const CdfGrid &g, int j,
                                             int tb, const NumericsProfile &np, uint4 *__restrict__ item, uint16_t *stage,
                                             int lane) {
    const int last = g.Lp - 1;
    const int t = tb * 32 + lane;
    const long long i = (long long)j + (long long)t * dg.S;
    GmmChannel ch;
#pragma unroll
    for (int m = 0; m < kM; ++m) { ch.sigma[m] = 1.f; ch.mu[m] = 0.f; ch.w[m] = 0.2f; ch.rinv[m] = 1.f; }
    ch.fast = 1;
    int base = 0;
    if (kPipe && clr >= 1) {
        // Wait for the previous channel's symbols with ONE lane on the item's last symbol (the consumer stores the 32
        // symbols of an item with one instruction; a Co symbol implies the Y symbols before it).  Thousands of producer
        // warps wait here at any time: polled by every lane, their loads alone are a large share of the L2 traffic.
        if (lane == 0) {
            const long long steps = (dg.n_sym - j + dg.S - 1) / dg.S;
            const long long i_last = (long long)j + (min((long long)tb * 32 + 31, steps - 1)) * dg.S;
            const int16_t *p = syms + (size_t)(clr - 1) * sym_cap + i_last;
            for (uint32_t polls = 0; ld_relaxed_s16(p) == (int)kSentinel; ++polls) {
                if (give_up(polls)) break;
                __nanosleep(500);
            }
        }
        __syncwarp();
    }
    if (i < dg.n_sym) {
        const int r = (int)(i / dg.crop_w), c = (int)(i - (long long)r * dg.crop_w);
        const size_t pidx = (size_t)r * dg.Ws + c;
        int y0 = 0, y1 = 0;
        if (clr >= 1) y0 = read_symbol<kPipe>(syms + i) + lo[0];
        if (clr == 2) y1 = read_symbol<kPipe>(syms + sym_cap + i) + lo[1];
        load_channel(pp, P, pidx, clr, y0, y1, np, ch);
        float mean = 0.f;
#pragma unroll
        for (int m = 0; m < kM; ++m) mean = fmaf(ch.w[m], ch.mu[m], mean);
        const int kc = __float2int_rn(mean * 255.0f) - g.min_val;
        base = min(max(kc - kWin / 2, 0), max(last - kWin, 0));
    }
    const int nv = min(32, (dg.n_sym - j + dg.S - 1) / dg.S - tb * 32);   // valid steps of this item (warp-uniform)
    // the prepared channel of every step goes through shared memory: 6 broadcast 16-byte reads per step
    float4 *prm = reinterpret_cast<float4 *>(stage + 32 * 34);            // [32 steps][6 x float4]
    {
        float4 *mine = prm + lane * 6;
        mine[0] = make_float4(ch.sigma[0], ch.sigma[1], ch.sigma[2], ch.sigma[3]);
        mine[1] = make_float4(ch.sigma[4], ch.mu[0], ch.mu[1], ch.mu[2]);
        mine[2] = make_float4(ch.mu[3], ch.mu[4], ch.w[0], ch.w[1]);
        mine[3] = make_float4(ch.w[2], ch.w[3], ch.w[4], ch.rinv[0]);
        mine[4] = make_float4(ch.rinv[1], ch.rinv[2], ch.rinv[3], ch.rinv[4]);
        mine[5] = make_float4(__int_as_float(base), __int_as_float(ch.fast), 0.f, 0.f);
    }
    __syncwarp();
    uint16_t *my = stage + lane * 34;     // 17-word pitch: conflict-free writes and reads
#pragma unroll 1
    for (int s = 0; s < nv; ++s) {
        const float4 *src = prm + s * 6;
        const float4 v0 = src[0], v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4], v5 = src[5];
        GmmChannel b;
        b.sigma[0] = v0.x; b.sigma[1] = v0.y; b.sigma[2] = v0.z; b.sigma[3] = v0.w; b.sigma[4] = v1.x;
        b.mu[0] = v1.y; b.mu[1] = v1.z; b.mu[2] = v1.w; b.mu[3] = v2.x; b.mu[4] = v2.y;
        b.w[0] = v2.z; b.w[1] = v2.w; b.w[2] = v3.x; b.w[3] = v3.y; b.w[4] = v3.z;
        b.rinv[0] = v3.w; b.rinv[1] = v4.x; b.rinv[2] = v4.y; b.rinv[3] = v4.z; b.rinv[4] = v4.w;
        const int bs = __float_as_int(v5.x);
        b.fast = __float_as_int(v5.y);
        const int k = bs + lane;
        // every lane evaluates an entry (the warp-uniform erfc skip votes over all 32 lanes); lanes at or
        // beyond the alphabet's end evaluate a clamped index and store 0 (= 2^16 mod 2^16)
        const uint32_t qe = b.fast ? cdf_q<true, true>(b, g, min(k, last - 1), np) : cdf_q<true, false>(b, g, min(k, last - 1), np);
        my[s] = (uint16_t)(k < last ? qe : 0u);
    }
    __syncwarp();
    const uint32_t *mw = reinterpret_cast<const uint32_t *>(my);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        uint4 o;
        o.x = mw[4 * v + 0]; o.y = mw[4 * v + 1]; o.z = mw[4 * v + 2]; o.w = mw[4 * v + 3];
        if (kPipe) __stcg(item + v * 32 + lane, o); else item[v * 32 + lane] = o;
    }
    unsigned short *bases = reinterpret_cast<unsigned short *>(item + 128);
    if (kPipe) __stcg(bases + lane, (unsigned short)base); else bases[lane] = (unsigned short)base;
    if (kPipe) {      // few, long chains: a symbol outside its window should not cost the consumer a reload of 60 planes
        float4 *dst = reinterpret_cast<float4 *>(item + 132);
#pragma unroll
        for (int k = 0; k < 6; ++k) __stcg(dst + k * 32 + lane, prm[k * 32 + lane]);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Consumer: the serial arithmetic decoder of one chain over pre-computed windows.
// ------------------------------------------------------------------------------------------
// Slow path: the symbol is outside the window.  Full analytic search (identical result to
// torchac's binary search over the dense row).
// Returns c_low | (c_high - 1) << 16 | symbol << 32.
__device__ __noinline__ uint64_t slow_symbol(const float *__restrict__ pp, const int16_t *syms, size_t sym_cap, size_t P,
                                             int crop_w, int Ws, long long i, int clr, int lo0, int lo1, CdfGrid g,
                                             NumericsProfile np, uint32_t low, uint32_t high, uint32_t value, int lane,
                                             int pipe, int first_base) {
    if (lane == 0) atomicAdd(&g_decode_stats[0], 1ull);
    const int r = (int)(i / crop_w), c = (int)(i - (long long)r * crop_w);
    const size_t pidx = (size_t)r * Ws + c;
    int y0 = 0, y1 = 0;
    if (clr >= 1) y0 = (pipe ? ld_relaxed_s16(syms + i) : (int)syms[i]) + lo0;
    if (clr == 2) y1 = (pipe ? ld_relaxed_s16(syms + sym_cap + i) : (int)syms[sym_cap + i]) + lo1;
    WarpParams wp;
    wp.load(pp, P, pidx, lane);
    GmmChannel ch;
    warp_channel(wp, clr, y0, y1, np, ch);
    const uint64_t span = (uint64_t)high - (uint64_t)low + 1ull;
    const uint64_t num = (((uint64_t)value - (uint64_t)low + 1ull) << 16) - 1ull;
    const uint32_t target = (uint32_t)(num / span) & 0xFFFFu;
    uint32_t c_low, c_high;
    const int sym = warp_search(ch, g, target, np, lane, c_low, c_high, first_base);
    return (uint64_t)c_low | ((uint64_t)(c_high - 1u) << 16) | ((uint64_t)(uint32_t)sym << 32);
}

// The same from the prepared channel the producer left beside the item (piped schedules): six broadcast loads
// instead of 60 strided planes, two symbols, the coupling and the weight normalisation.
__device__ __noinline__ uint64_t slow_symbol_prepared(const float4 *__restrict__ prm, CdfGrid g, NumericsProfile np, uint32_t low,
                                                      uint32_t high, uint32_t value, int lane, int first_base) {
    if (lane == 0) atomicAdd(&g_decode_stats[0], 1ull);
    const float4 v0 = __ldcg(prm), v1 = __ldcg(prm + 1), v2 = __ldcg(prm + 2), v3 = __ldcg(prm + 3), v4 = __ldcg(prm + 4), v5 = __ldcg(prm + 5);
    GmmChannel b;
    b.sigma[0] = v0.x; b.sigma[1] = v0.y; b.sigma[2] = v0.z; b.sigma[3] = v0.w; b.sigma[4] = v1.x;
    b.mu[0] = v1.y; b.mu[1] = v1.z; b.mu[2] = v1.w; b.mu[3] = v2.x; b.mu[4] = v2.y;
    b.w[0] = v2.z; b.w[1] = v2.w; b.w[2] = v3.x; b.w[3] = v3.y; b.w[4] = v3.z;
    b.rinv[0] = v3.w; b.rinv[1] = v4.x; b.rinv[2] = v4.y; b.rinv[3] = v4.z; b.rinv[4] = v4.w;
    b.fast = __float_as_int(v5.y);
    const uint64_t span = (uint64_t)high - (uint64_t)low + 1ull;
    const uint64_t num = (((uint64_t)value - (uint64_t)low + 1ull) << 16) - 1ull;
    const uint32_t target = (uint32_t)(num / span) & 0xFFFFu;
    uint32_t c_low, c_high;
    const int sym = warp_search(b, g, target, np, lane, c_low, c_high, first_base);
    return (uint64_t)c_low | ((uint64_t)(c_high - 1u) << 16) | ((uint64_t)(uint32_t)sym << 32);
}

// Serial coder state of one chain: torchac's (low, high, value) registers and the stream position.
// The next 128 stream bits are rebuilt from memory (L1 hits) at the start of every 8-step chunk, so
// inside a chunk there is no memory access, no refill and no branch.
struct ChainCoder {
    uint32_t low, high, value;
    uint32_t pos;                  // stream bits already shifted into `value`, counted from w
    uint32_t b0, b1, b2, b3;       // the stream bits after pos, left aligned (transient: load_window)
    int avail;                     // valid bits in b0:b1:b2:b3
    const uint32_t *w;             // 4-byte aligned base (<= stream start)
    uint32_t hi_byte;              // end of the stream relative to w, in bytes

    // big-endian word i of the stream; torchac reads zeros past the end
    __device__ __forceinline__ uint32_t word(uint32_t i) const {
        const uint32_t p0 = i * 4u;
        if (p0 >= hi_byte) return 0u;
        uint32_t v = __byte_perm(__ldg(w + i), 0, 0x0123);       // first stream byte in the most significant position
        if (p0 + 4u > hi_byte) v &= 0xFFFFFFFFu << (8u * (p0 + 4u - hi_byte));
        return v;
    }
    __device__ __forceinline__ void load_window() {
        const uint32_t wi = pos >> 5, r = pos & 31u;
        uint32_t W0, W1, W2, W3, W4;
        if ((wi + 5u) * 4u <= hi_byte) {        // all five words inside the stream (uniform; false only at its end)
            W0 = __byte_perm(__ldg(w + wi), 0, 0x0123);
            W1 = __byte_perm(__ldg(w + wi + 1), 0, 0x0123);
            W2 = __byte_perm(__ldg(w + wi + 2), 0, 0x0123);
            W3 = __byte_perm(__ldg(w + wi + 3), 0, 0x0123);
            W4 = __byte_perm(__ldg(w + wi + 4), 0, 0x0123);
            if ((wi + 70u) * 4u <= hi_byte) asm volatile("prefetch.global.L1 [%0];" ::"l"(w + wi + 64));
        } else {
            W0 = word(wi); W1 = word(wi + 1); W2 = word(wi + 2); W3 = word(wi + 3); W4 = word(wi + 4);
        }
        b0 = __funnelshift_l(W1, W0, r);
        b1 = __funnelshift_l(W2, W1, r);
        b2 = __funnelshift_l(W3, W2, r);
        b3 = __funnelshift_l(W4, W3, r);
        avail = 128;
    }
    __device__ __forceinline__ void shift(int sh) {     // 0 <= sh <= 31
        b0 = __funnelshift_l(b1, b0, sh);
        b1 = __funnelshift_l(b2, b1, sh);
        b2 = __funnelshift_l(b3, b2, sh);
        b3 <<= sh;
        avail -= sh;
    }
    __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t n) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
        w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        const uint32_t lo_byte = (uint32_t)(a & 3);
        hi_byte = lo_byte + n;
        pos = 8u * lo_byte;
        load_window();
        low = 0; high = 0xFFFFFFFFu;
        value = b0;
        pos += 32u;
    }
};

// Build-time variants (tools/step_probe.cu measures them on the GPU):
//   LLICTI_CLZ_I2F  1  count leading zeros through the exponent of a round-toward-zero int -> float
//                      conversion (I2F: ~10 cycles) instead of FLO (~22 cycles)
#ifndef LLICTI_CLZ_I2F
#define LLICTI_CLZ_I2F 1
#endif
#ifndef LLICTI_PROBE_X
#define LLICTI_PROBE_X 0     // measurement-only knobs of tools/step_probe.cu (never set in the product build)
#endif

// clz of a non-zero word (the argument below is non-zero whenever nl < nh; other lanes' results are never selected)
__device__ __forceinline__ int clz_nz(uint32_t x) {
#if LLICTI_CLZ_I2F
    return 158 - (int)(__float_as_uint(__uint2float_rz(x)) >> 23);
#else
    return __clz(x);
#endif
}

// torchac's interval update and renormalisation for the symbol with bounds (c_low, c_high), as a
// pure function of the coder registers.  The bit-serial E1/E2/E3 loop shifts out
//   n = clz(nl ^ nh)                          equal leading bits (E1/E2 run), then
//   k = the run of (1, 0) bit pairs after the first differing pair  (E3 / underflow run).
// With d = nl ^ nh and m = nl & ~nh (ones at the (1, 0) pairs; the run occupies n+1 .. n+k),
// d & ~(m << 1) has its first one exactly at n + k, so ONE clz gives the total shift sh = n + k.
// After the shift the top pair is (0, 1) whether or not an underflow happened; it happened iff
// the top bit of nl << sh is set, which is also the bit torchac flips in `value`.
struct NextState { uint32_t low, high, value; int sh; uint32_t nl, nh; };
__device__ __forceinline__ NextState renormalise(uint32_t nl, uint32_t nh, uint32_t value, uint32_t next_bits) {
    NextState o;
    o.nl = nl; o.nh = nh;
    o.sh = clz_nz((nl ^ nh) & ~((nl & ~nh) << 1));                      // <= 31 whenever nl < nh
    const uint32_t ls = nl << o.sh;
    o.low = ls & 0x7FFFFFFFu;
    o.high = (nh << o.sh) | ~(0xFFFFFFFFu << o.sh) | 0x80000000u;
    o.value = __funnelshift_l(next_bits, value, o.sh) ^ (ls & 0x80000000u);
    return o;
}
// exact for every state (64-bit products, as torchac)
__device__ __forceinline__ NextState next_state(uint32_t low, uint32_t high, uint32_t value, uint32_t next_bits,
                                                uint32_t c_low, uint32_t c_high) {
    const uint32_t sm1 = high - low;                                    // span - 1; span * c = sm1 * c + c
    const uint32_t nl = low + (uint32_t)(((uint64_t)sm1 * c_low + c_low) >> 16);
    const uint32_t nh = (low - 1u) + (uint32_t)(((uint64_t)sm1 * c_high + c_high) >> 16);
    return renormalise(nl, nh, value, next_bits);
}

template <bool kPipe>
__device__ __forceinline__ uint4 load_chunk(const uint4 *p) { return kPipe ? __ldcg(p) : *p; }
template <bool kPipe>
__device__ __forceinline__ int load_base(const uint4 *item, int lane) {
    const unsigned short *b = reinterpret_cast<const unsigned short *>(item + 128) + lane;
    return (int)(kPipe ? __ldcg(b) : *b);
}

// entry of step e (0..7) of a chunk word: as the 16-bit value, and shifted to the top half of a word
__device__ __forceinline__ uint32_t chunk_entry(const uint4 &q, int e) {
    const uint32_t w = (e >> 1) == 0 ? q.x : (e >> 1) == 1 ? q.y : (e >> 1) == 2 ? q.z : q.w;
    return (e & 1) ? (w >> 16) : (w & 0xFFFFu);
}
__device__ __forceinline__ uint32_t chunk_entry_hi(const uint4 &q, int e) {
    const uint32_t w = (e >> 1) == 0 ? q.x : (e >> 1) == 1 ? q.y : (e >> 1) == 2 ? q.z : q.w;
    return (e & 1) ? (w & 0xFFFF0000u) : (w << 16);
}

struct ChainCtx {       // what the (rare) slow path needs
    const float *pp;
    const int16_t *syms;
    size_t sym_cap, P;
    int crop_w, Ws, clr, lo0, lo1, S;
};

// One symbol.  Lane l holds table entry q(base + l) of this symbol's window (cl16 = q << 16) and evaluates,
// for the candidate "symbol = base + l", the complete next coder state; which candidate is right is
//   low + ((span * q(s)) >> 16) <= value          (<=> q(s) <= floor(((value-low+1) 2^16 - 1) / span),
// torchac's search key), so a ballot picks the lane and four shuffles fetch its state.  The only
// work after the ballot is the selection.
//
// Fast variant: no branch at all, 32-bit arithmetic.  (span * q) >> 16 is the high word of
// span * (q << 16) whenever span < 2^32; an upper bound of 2^16 (stored as 0: the top symbol of the
// alphabet) gives `span` itself.  A full-range state (span = 2^32), a symbol outside its window or a
// window of stream bits that ran low only raise `bad`; the caller validates once per chunk and
// redoes a bad chunk with the careful variant from a snapshot (about 1 chunk in 100).
// Returns the window-relative symbol.
__device__ __forceinline__ uint32_t decode_step_fast(ChainCoder &cc, uint32_t cl16, uint32_t vmask, uint32_t &bad) {
    const uint32_t ch16 = __shfl_down_sync(kFull, cl16, 1);              // data only, off the serial chain
    const uint32_t span = cc.high - cc.low + 1u;                         // 0: the full range 2^32
    const uint32_t nl = cc.low + __umulhi(span, cl16);
    const uint32_t nhp1 = cc.low + (ch16 == 0u ? span : __umulhi(span, ch16));   // nh + 1 = nl of the next candidate
    const NextState ns = renormalise(nl, nhp1 - 1u, cc.value, cc.b0);
    // beyond the window: only possible for the last candidate (for the others it contradicts the ballot)
    const uint32_t shw = (uint32_t)ns.sh | (cc.value >= nhp1 ? 0x100u : 0u);
#if LLICTI_PROBE_X == 4                          // redux.max instead of ballot + popc
    const uint32_t li = __reduce_max_sync(kFull, (cc.value >= nl && ((vmask >> (threadIdx.x & 31)) & 1u)) ? (threadIdx.x & 31) + 1u : 0u) - 1u;
#else
    const uint32_t li = (uint32_t)__popc(__ballot_sync(kFull, cc.value >= nl) & vmask) - 1u;
#endif
    bad |= (span == 0u ? 1u : 0u) | (cc.avail < 32 ? 1u : 0u);
    cc.low = __shfl_sync(kFull, ns.low, li);
    cc.high = __shfl_sync(kFull, ns.high, li);
    cc.value = __shfl_sync(kFull, ns.value, li);
    const uint32_t w = __shfl_sync(kFull, shw, li);
    bad |= (li >> 31) | (w >> 8);                                        // li = -1: below the window
    cc.shift((int)(w & 31u));
    return li << 8;
}

// Careful variant: exact in every state; a symbol outside its window takes the full analytic search.
// Returns the symbol (alphabet index).
template <bool kPipe>
__device__ __forceinline__ int decode_step_careful(ChainCoder &cc, uint32_t slot, int base, uint32_t vmask, long long i,
                                               const ChainCtx &cx, const CdfGrid &g, const NumericsProfile &np, int lane,
                                               const float4 *prm) {
    cc.load_window();
    const uint32_t up = __shfl_down_sync(kFull, slot, 1);
    const uint32_t c_high = up == 0u ? 0x10000u : up;
    const NextState ns = next_state(cc.low, cc.high, cc.value, cc.b0, slot, c_high);
    const uint32_t li = (uint32_t)__popc(__ballot_sync(kFull, cc.value >= ns.nl) & vmask) - 1u;
    const uint32_t nh_sel = __shfl_sync(kFull, ns.nh, li);
    int sym, sh;
    if ((li >> 31) != 0u || cc.value > nh_sel) {
        const int first_base = (li >> 31) != 0u ? max(base - 31, 0) : base + kWin;      // right below / above the window that missed
        const uint64_t pk = kPipe ? slow_symbol_prepared(prm, g, np, cc.low, cc.high, cc.value, lane, first_base)
                                  : slow_symbol(cx.pp, cx.syms, cx.sym_cap, cx.P, cx.crop_w, cx.Ws, i, cx.clr, cx.lo0, cx.lo1, g, np,
                                                cc.low, cc.high, cc.value, lane, 0, first_base);
        const NextState s2 = next_state(cc.low, cc.high, cc.value, cc.b0, (uint32_t)pk & 0xFFFFu,
                                        (((uint32_t)pk >> 16) & 0xFFFFu) + 1u);
        cc.low = s2.low; cc.high = s2.high; cc.value = s2.value; sh = s2.sh;
        sym = (int)(pk >> 32);
    } else {
        cc.low = __shfl_sync(kFull, ns.low, li);
        cc.high = __shfl_sync(kFull, ns.high, li);
        cc.value = __shfl_sync(kFull, ns.value, li);
        sh = __shfl_sync(kFull, ns.sh, li);
        sym = base + (int)li;
    }
    cc.pos += (uint32_t)sh;     // (torchac does not update after the last symbol; the state is dead by then)
    return sym;
}

// Coder registers of a chain between two launches of the wavefront schedule (the chain is decoded a
// strip of items at a time).
struct ChainState { uint32_t low, high, value, pos; };

// `out` = compact symbol array of this (image, channel) in coding order; chain j writes j, j+S, ...
// Items [it_begin, it_end) of the chain are decoded; it_begin > 0 resumes from *state, it_end short
// of the chain's end saves to it.  `li_buf` = 32 words of shared memory of this warp: the window-relative
// symbols of the item in flight; lane t adds the base of step t and stores the symbol.
template <bool kPipe>
__device__ __forceinline__ void consume_chain(const ChainCtx &cx, int n_sym, const CdfGrid &g, int j,
                                              const NumericsProfile &np, int16_t *out, const uint4 *__restrict__ items,
                                              uint32_t *flags, const uint8_t *__restrict__ stream, uint32_t stream_len,
                                              int lane, int *li_buf, int it_begin = 0, int it_end = 0x7FFFFFFF,
                                              ChainState *state = nullptr) {

    const int S = cx.S;
    const int n_steps = (n_sym - j + S - 1) / S;
    if (n_steps <= 0) return;
    const int n_items_all = (n_steps + 31) >> 5;      // an item = 4 chunks of 8 steps (one uint4 per lane each)
    const int n_items = min(n_items_all, it_end);     // decode up to here in this call
    if (it_begin >= n_items) return;
    const int n_full = n_steps >> 3, tail = n_steps & 7;
    const int last = g.Lp - 1;
    const uint32_t vmask = last >= kWin ? 0x7FFFFFFFu : (1u << last) - 1u;    // candidate lanes that are symbols
    unsigned long long polls = 0, redone = 0;
    long long waited = 0, redo_cycles = 0;
    const long long t_begin = kPipe ? clock64() : 0;

    ChainCoder cc;
    cc.init(stream, stream_len);
    if (it_begin > 0) {
        const ChainState s = *state;
        cc.low = s.low; cc.high = s.high; cc.value = s.value; cc.pos = s.pos;
    }
    long long i = j + (long long)it_begin * 32 * S;

    if (kPipe && !wait_flag(flags + it_begin, polls, waited)) return;      // abandoned launch (g_abort)
    const long long first_wait = waited;
    const uint4 *src = items + lane;
    const uint4 *first = src + (size_t)it_begin * kItemU4;
    uint4 q0 = load_chunk<kPipe>(first), q1 = load_chunk<kPipe>(first + 32), q2 = load_chunk<kPipe>(first + 64),
          q3 = load_chunk<kPipe>(first + 96);
    int base_cur = load_base<kPipe>(items + (size_t)it_begin * kItemU4, lane);
    uint4 n0 = q0, n1 = q1, n2 = q2, n3 = q3;
    int base_nxt = base_cur;
    uint32_t f_next = kPipe && it_begin + 1 < n_items ? ld_relaxed_u32(flags + it_begin + 1) : 1u;   // looked at one item later

    for (int it = it_begin; it < n_items; ++it) {
        // the next item is fetched while this one is decoded (32 steps of distance)
        bool have_next = it + 1 >= n_items;
        uint32_t f_next2 = 1u;
        if (it + 1 < n_items) {
            if (!kPipe || f_next != 0u) {
                const uint4 *nx = src + (size_t)(it + 1) * kItemU4;
                n0 = load_chunk<kPipe>(nx); n1 = load_chunk<kPipe>(nx + 32); n2 = load_chunk<kPipe>(nx + 64); n3 = load_chunk<kPipe>(nx + 96);
                base_nxt = load_base<kPipe>(items + (size_t)(it + 1) * kItemU4, lane);
                have_next = true;
            }
            if (kPipe && it + 2 < n_items) f_next2 = ld_relaxed_u32(flags + it + 2);
        }
        const int full_here = min(4, n_full - it * 4);
        const float4 *item_prm = reinterpret_cast<const float4 *>(items + (size_t)it * kItemU4 + 132);   // prepared channels (piped schedules)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (v >= full_here) break;
            const uint4 q = v == 0 ? q0 : v == 1 ? q1 : v == 2 ? q2 : q3;
            cc.load_window();
            const uint32_t s_low = cc.low, s_high = cc.high, s_value = cc.value;
            uint32_t bad = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) li_buf[8 * v + e] = (int)decode_step_fast(cc, chunk_entry_hi(q, e), vmask, bad);
            if (__builtin_expect(bad != 0u, 0)) {
                cc.low = s_low; cc.high = s_high; cc.value = s_value;      // cc.pos is only advanced below
                ++redone;
                const long long t_redo = kPipe ? clock64() : 0;
                // step by step over one window of stream bits: the fast step, validated at once; only the step that
                // fails is repeated carefully (cc.pos moves to that step first, the window is rebuilt after it)
                cc.load_window();
#pragma unroll 1
                for (int e = 0; e < 8; ++e) {
                    const uint32_t r_low = cc.low, r_high = cc.high, r_value = cc.value, r0 = cc.b0, r1 = cc.b1, r2 = cc.b2, r3 = cc.b3;
                    const int r_avail = cc.avail;
                    uint32_t bad1 = 0;
                    const uint32_t wfast = decode_step_fast(cc, chunk_entry_hi(q, e), vmask, bad1);
                    if (bad1 == 0u) {
                        li_buf[8 * v + e] = (int)wfast;
                    } else {
                        cc.low = r_low; cc.high = r_high; cc.value = r_value; cc.b0 = r0; cc.b1 = r1; cc.b2 = r2; cc.b3 = r3;
                        cc.pos += (uint32_t)(128 - r_avail);
                        const int base = __shfl_sync(kFull, base_cur, 8 * v + e);
                        const int sy = decode_step_careful<kPipe>(cc, chunk_entry(q, e), base, vmask, i + (long long)(8 * v + e) * S, cx, g, np, lane,
                                                                  item_prm + (8 * v + e) * 6);
                        li_buf[8 * v + e] = (int)(((uint32_t)(sy - base) & 0xFFFu) << 8);
                        cc.load_window();
                    }
                }
                if (kPipe) redo_cycles += clock64() - t_redo;
            }
            cc.pos += (uint32_t)(128 - cc.avail);
        }
        if (it == n_items_all - 1 && tail) {
            const int v = max(full_here, 0);
            const uint4 q = v == 0 ? q0 : v == 1 ? q1 : v == 2 ? q2 : q3;
#pragma unroll 1
            for (int e = 0; e < tail; ++e) {
                const int base = __shfl_sync(kFull, base_cur, 8 * v + e);
                const int sy = decode_step_careful<kPipe>(cc, chunk_entry(q, e), base, vmask, i + (long long)(8 * v + e) * S, cx, g, np, lane,
                                                                  item_prm + (8 * v + e) * 6);
                li_buf[8 * v + e] = (int)(((uint32_t)(sy - base) & 0xFFFu) << 8);
            }
        }
        // lane t owns step t of the item: symbol = window base + window-relative symbol
        __syncwarp();
        if (lane < min(32, n_steps - it * 32)) {
            const int sy = base_cur + (((int)((uint32_t)li_buf[lane] << 12)) >> 20);     // bits 8..19: window-relative symbol, signed
            int16_t *dst = out + i + (long long)lane * S;
            if (kPipe) st_relaxed_s16(dst, sy); else *dst = (int16_t)sy;
        }
        __syncwarp();
        i += 32ll * S;
        if (!have_next) {
            if (!wait_flag(flags + it + 1, polls, waited)) return;          // abandoned launch (g_abort)
            const uint4 *nx = src + (size_t)(it + 1) * kItemU4;
            n0 = load_chunk<kPipe>(nx); n1 = load_chunk<kPipe>(nx + 32); n2 = load_chunk<kPipe>(nx + 64); n3 = load_chunk<kPipe>(nx + 96);
            base_nxt = load_base<kPipe>(items + (size_t)(it + 1) * kItemU4, lane);
        }
        q0 = n0; q1 = n1; q2 = n2; q3 = n3;
        base_cur = base_nxt;
        f_next = f_next2;
    }
    if (state != nullptr && n_items < n_items_all && lane == 0) {
        ChainState s;
        s.low = cc.low; s.high = cc.high; s.value = cc.value; s.pos = cc.pos;
        *state = s;
    }
    if (lane == 0) {
        if (kPipe && polls) atomicAdd(&g_decode_stats[1], polls);
        if (kPipe) {
            atomicAdd(&g_decode_stats[4], (unsigned long long)waited);
            atomicAdd(&g_decode_stats[5], (unsigned long long)(clock64() - t_begin));
            atomicAdd(&g_decode_stats[6], 1ull);
            atomicMax(&g_decode_stats[2], (unsigned long long)(clock64() - t_begin));
            atomicAdd(&g_wave_dbg[1], (unsigned long long)(clock64() - t_begin));
            atomicAdd(&g_wave_dbg[2], (unsigned long long)waited);
            atomicAdd(&g_wave_dbg[3], (unsigned long long)redo_cycles);
            atomicAdd(&g_wave_dbg[4 + min(cx.clr, 2)], (unsigned long long)waited);     // waiting per colour channel
            atomicAdd(&g_decode_stats[7], (unsigned long long)redo_cycles);
        }
        if (redone) atomicAdd(&g_decode_stats[3], redone);
    }
}

// Compact symbols of a decoded band -> centred samples in the planes, with the replicate padding
// of the short phases (_pad_decoded_tensor, LLICTI_nets.py:512-530).
__global__ void __launch_bounds__(256)
scatter_band_kernel(const int16_t *__restrict__ syms, size_t sym_cap, int16_t *__restrict__ planes,
                    const int32_t *__restrict__ minmax, DecodeGeom dg, int n, int sym0, int sym1) {
    const int img = blockIdx.y / 3, clr = blockIdx.y - 3 * img;
    const int i = sym0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sym1) return;
    const bool rep_w = dg.padW && (dg.band == 0 || dg.band == 1);
    const bool rep_h = dg.padH && (dg.band == 0 || dg.band == 2);
    const size_t P = (size_t)dg.Hs * dg.Ws;
    const int32_t *mm = minmax + img * 4;
    const int lo = clr == 0 ? -127 : clr == 1 ? mm[0] : mm[1];
    const int r = i / dg.crop_w, c = i - r * dg.crop_w;
    const int16_t v = (int16_t)((int)syms[((size_t)img * 3 + clr) * sym_cap + i] + lo);
    int16_t *dst = planes + (size_t)img * 12 * P + (size_t)(3 * (dg.band + 1) + clr) * P + (size_t)r * dg.Ws + c;
    dst[0] = v;
    const bool last_c = rep_w && c == dg.crop_w - 1, last_r = rep_h && r == dg.crop_h - 1;
    if (last_c) dst[1] = v;
    if (last_r) dst[dg.Ws] = v;
    if (last_c && last_r) dst[dg.Ws + 1] = v;
}

// Alphabet of every colour channel of an image: Y is fixed, Co / Cg come from the header's min / max.
// (Selected by value: a CdfGrid array indexed by a run-time channel would live in local memory.)
static __device__ __forceinline__ CdfGrid band_grids(const int32_t *mm, int clr, int (&lo)[3]) {
    lo[0] = -127; lo[1] = mm[0]; lo[2] = mm[1];
    return clr == 0 ? make_grid(-127, 128) : clr == 1 ? make_grid(mm[0], mm[2]) : make_grid(mm[1], mm[3]);
}

// ---- split schedule ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
window_kernel(const float *__restrict__ params, const int16_t *__restrict__ syms, size_t sym_cap,
              const int32_t *__restrict__ minmax, DecodeGeom dg, int clr, NumericsProfile np, uint4 *__restrict__ items,
              int n) {
    __shared__ __align__(16) uint16_t stage[4][kStageU16];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long per_img = (long long)dg.S * dg.items_per_chain;
    const long long total = per_img * n;
    for (long long w = (long long)blockIdx.x * 4 + wib; w < total; w += (long long)gridDim.x * 4) {
        const int img = (int)(w / per_img);
        const long long rem = w - (long long)img * per_img;
        // chain-minor order: neighbouring warps work on neighbouring chains of the same step block,
        // so their (strided) parameter reads share sectors
        const int tb = (int)(rem / dg.S), j = (int)(rem % dg.S);
        if ((long long)j + (long long)tb * 32 * dg.S >= dg.n_sym) continue;
        const size_t P = (size_t)dg.Hs * dg.Ws;
        int lo[3];
        const CdfGrid g = band_grids(minmax + img * 4, clr, lo);
        uint4 *item = items + ((((size_t)img * 3 + clr) * dg.S + j) * dg.items_per_chain + tb) * kItemU4;
        produce_item<false>(params + (size_t)img * kParamCh * P, syms + (size_t)img * 3 * sym_cap, sym_cap, P, dg, clr, lo,
                            g, j, tb, np, item, stage[wib], lane);
    }
}

__global__ void __launch_bounds__(128)
consume_kernel(const float *__restrict__ params, int16_t *__restrict__ syms, size_t sym_cap,
               const int32_t *__restrict__ minmax, DecodeGeom dg, int clr, NumericsProfile np,
               const uint4 *__restrict__ items, const uint8_t *__restrict__ blob, const uint64_t *__restrict__ suboff,
               const uint32_t *__restrict__ sublen, int total_sub, int n) {
    __shared__ __align__(16) int li_buf[4][kLiBuf];
    const int lane = threadIdx.x & 31;
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= (long long)n * dg.S) return;
    const int img = (int)(w / dg.S), j = (int)(w - (long long)img * dg.S);
    const size_t P = (size_t)dg.Hs * dg.Ws;
    int lo[3];
    const CdfGrid g = band_grids(minmax + img * 4, clr, lo);
    int16_t *isyms = syms + (size_t)img * 3 * sym_cap;
    const ChainCtx cx = {params + (size_t)img * kParamCh * P, isyms, sym_cap, P, dg.crop_w, dg.Ws, clr, lo[0], lo[1], dg.S};
    const size_t e = (size_t)img * total_sub + dg.sub_first[clr] + j;
    const uint4 *chain_items = items + (((size_t)img * 3 + clr) * dg.S + j) * dg.items_per_chain * kItemU4;
    consume_chain<false>(cx, dg.n_sym, g, j, np, isyms + (size_t)clr * sym_cap, chain_items, nullptr, blob + suboff[e],
                         sublen[e], lane, li_buf[threadIdx.x >> 5]);
}

// ---- piped schedule (S == 1): one grid of one-warp CTAs, all co-resident ----------------------
// Roles are claimed at run time so that the serial consumer warps get SMs of their own: the first
// cons_per_sm CTAs to arrive on an SM draw consumer tickets (one chain each, at most one per
// scheduler); once the tickets are gone the remaining SMs are producer SMs, and CTAs that land on
// a consumer SM without a ticket leave.  Producers draw window items through per-channel tickets
// in (step block, image) order: Y items wait for nothing; a Co (Cg) item waits for the Y (and Co)
// symbols of its 32 positions, i.e. for consumers that only depend on items earlier in the same
// orders -- the holder of the earliest unfinished ticket can always run, so the grid cannot
// deadlock as long as it is co-resident.
constexpr int kConsPerSmMax = 4;
constexpr int kCtlWords = 1024;     // [0,256) arrivals per SM, [256,512) role per SM, 512 consumer / 513 producer / 514.. item tickets

__global__ void __launch_bounds__(32)
decode_band_pipe_kernel(const float *__restrict__ params, int16_t *syms, size_t sym_cap, const int32_t *__restrict__ minmax,
                        DecodeGeom dg, NumericsProfile np, uint4 *items, uint32_t *ctl, uint32_t *flags, int cons_per_sm,
                        const uint8_t *__restrict__ blob, const uint64_t *__restrict__ suboff,
                        const uint32_t *__restrict__ sublen, int total_sub, int n) {
    __shared__ __align__(16) uint16_t stage[kStageU16];
    __shared__ __align__(16) int li_buf[kLiBuf];
    const int lane = threadIdx.x;
    const size_t P = (size_t)dg.Hs * dg.Ws;
    const int n_cons = 3 * n;
    int lo[3];

    // ---- role ----
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smid &= 255u;
    int chain = -1, producer_id = -1;
    if (lane == 0) {
        const uint32_t slot = atomicAdd(&ctl[smid], 1u);
        if (slot < (uint32_t)cons_per_sm) {
            const uint32_t t = atomicAdd(&ctl[512], 1u);
            if (t < (uint32_t)n_cons) chain = (int)t;
            if (slot == 0) st_release_u32(&ctl[256 + smid], chain >= 0 ? 1u : 2u);
        }
        if (chain < 0) {
            uint32_t role;
            for (uint32_t spins = 0; (role = ld_relaxed_u32(&ctl[256 + smid])) == 0u; ++spins) {
                if (give_up(spins)) break;
                __nanosleep(100);
            }
            if (role == 2u) producer_id = (int)atomicAdd(&ctl[513], 1u);
        }
    }
    chain = __shfl_sync(kFull, chain, 0);
    producer_id = __shfl_sync(kFull, producer_id, 0);

    if (chain >= 0) {
        const int img = chain / 3, clr = chain - 3 * img;
        const CdfGrid g = band_grids(minmax + img * 4, clr, lo);
        int16_t *isyms = syms + (size_t)img * 3 * sym_cap;
        const ChainCtx cx = {params + (size_t)img * kParamCh * P, isyms, sym_cap, P, dg.crop_w, dg.Ws, clr, lo[0], lo[1], 1};
        const size_t e = (size_t)img * total_sub + dg.sub_first[clr];
        consume_chain<true>(cx, dg.n_sym, g, 0, np, isyms + (size_t)clr * sym_cap,
                            items + (size_t)chain * dg.items_per_chain * kItemU4, flags + (size_t)chain * dg.items_per_chain,
                            blob + suboff[e], sublen[e], lane, li_buf);
        return;
    }
    if (producer_id < 0) return;       // on a consumer SM without a ticket
    if (g_test_starve) return;

    const int r5 = producer_id % 5;    // Y : Co : Cg producers start 1 : 2 : 2, then help the other channels
    const int clr0 = r5 == 0 ? 0 : r5 <= 2 ? 1 : 2;
    const uint32_t total = (uint32_t)dg.items_per_chain * (uint32_t)n;
    for (int c = 0; c < 3; ++c) {
        const int clr = (clr0 + c) % 3;
        for (;;) {
            uint32_t w = 0;
            if (lane == 0) w = atomicAdd(&ctl[514 + clr], 1u);
            w = __shfl_sync(kFull, w, 0);
            if (w >= total || aborted()) break;
            const int tb = (int)(w / (uint32_t)n);
            const int img = (int)(w - (uint32_t)tb * (uint32_t)n);
            const int ch = img * 3 + clr;
            const CdfGrid g = band_grids(minmax + img * 4, clr, lo);
            uint4 *item = items + ((size_t)ch * dg.items_per_chain + tb) * kItemU4;
            produce_item<true>(params + (size_t)img * kParamCh * P, syms + (size_t)img * 3 * sym_cap, sym_cap, P, dg, clr, lo,
                               g, 0, tb, np, item, stage, lane);
            // every lane's stores happen before the flag store: warp barrier, then a cumulative fence
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                st_release_u32(flags + (size_t)ch * dg.items_per_chain + tb, 1u);
            }
        }
    }
}

// ---- wavefront schedule: the three bands of a scale decoded concurrently, a strip of rows apart ----
// Band b+1 of a scale needs the network outputs computed from band b, but only from rows i-2 .. i+2:
// once band b is one (slightly shifted) strip ahead, band b+1 can follow.  One launch of this kernel decodes one strip
// of up to three bands at once (the same consumer / producer roles as decode_band_pipe_kernel, with the
// coder registers of every chain carried from launch to launch in ChainState); between launches the
// host runs the CNN of the next strips and scatters the finished ones into the planes.  The serial
// chain of a scale shrinks from 3 * n_sym steps to about (1 + 2 / strips) * n_sym.
struct WaveBand {
    DecodeGeom dg;
    const float *params;     // [n][60][P] network outputs of this band
    int16_t *syms;           // [n][3][sym_cap]
    uint4 *items;            // [n][3][items_per_chain][128]
    uint32_t *flags;         // [n][3][items_per_chain]
    int it0, it1;            // items of every chain decoded by this launch (it0 == it1: band idle)
};
struct WaveArgs { WaveBand b[3]; };

// Consumers and producers of a wavefront step are two kernels running concurrently on two streams
// (they only talk through the flags / sentinels in global memory): the consumer kernel is nine
// one-warp CTAs per image with the ~150 registers the serial loop wants; the producer kernel is
// compiled on its own (64 registers), so three times as many producer warps fit on an SM as in the
// single-kernel form.  The producer grid is sized to leave room for the consumer CTAs on every SM
// (see launch_decode_scale_wave), so the consumers are resident whichever kernel starts first, and a
// producer only holds a ticket while it runs: the earliest unfinished item can always complete.
constexpr int kWaveChainsMax = 12;       // chains (warps) per consumer CTA, at most

__global__ void __launch_bounds__(32 * kWaveChainsMax)
wave_consume_kernel(WaveArgs wa, const int32_t *__restrict__ minmax, NumericsProfile np, size_t sym_cap, uint32_t *ctl,
                    const uint8_t *__restrict__ blob, const uint64_t *__restrict__ suboff,
                    const uint32_t *__restrict__ sublen, int total_sub, int n, ChainState *states) {
    // blockDim.x / 32 chains per CTA, spread over the four schedulers; the CTA claims its SM: producer CTAs
    // that share it leave, so that nothing competes with the serial chains for issue slots and L1
    __shared__ __align__(16) int li_buf[kWaveChainsMax][kLiBuf];
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        st_release_u32(&ctl[256 + (smid & 255u)], 1u);
        atomicAdd(&ctl[523], 1u);                       // consumer CTAs that are running
    }
    int act[3], n_act = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b)
        if (wa.b[b].it1 > wa.b[b].it0) act[n_act++] = b;
    const int chain = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chain >= 3 * n * n_act) return;
    const int slot = chain / (3 * n), rem = chain - slot * 3 * n;
    const int band = slot == 0 ? act[0] : slot == 1 ? act[1] : act[2];
    const WaveBand &wb = band == 0 ? wa.b[0] : band == 1 ? wa.b[1] : wa.b[2];
    const DecodeGeom dg = wb.dg;
    const int img = rem / 3, clr = rem - 3 * img;
    const size_t P = (size_t)dg.Hs * dg.Ws;
    int lo[3];
    const CdfGrid g = band_grids(minmax + img * 4, clr, lo);
    int16_t *isyms = wb.syms + (size_t)img * 3 * sym_cap;
    const ChainCtx cx = {wb.params + (size_t)img * kParamCh * P, isyms, sym_cap, P, dg.crop_w, dg.Ws, clr, lo[0], lo[1], 1};
    const size_t e = (size_t)img * total_sub + dg.sub_first[clr];
    const size_t ch = (size_t)img * 3 + clr;
    consume_chain<true>(cx, dg.n_sym, g, 0, np, isyms + (size_t)clr * sym_cap, wb.items + ch * dg.items_per_chain * kItemU4,
                        wb.flags + ch * dg.items_per_chain, blob + suboff[e], sublen[e], lane, li_buf[threadIdx.x >> 5], wb.it0, wb.it1,
                        states + ((size_t)img * 3 + band) * 3 + clr);
}

__global__ void __launch_bounds__(128, 8)
wave_produce_kernel(WaveArgs wa, const int32_t *__restrict__ minmax, NumericsProfile np, size_t sym_cap, uint32_t *ctl, int n,
                    int safe_ctas, int consumer_ctas, int share, int pattern) {
    __shared__ __align__(16) uint16_t stage[4][kStageU16];
    __shared__ uint32_t started;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // The first `safe_ctas` CTAs leave room on every SM for a consumer CTA.  The CTAs beyond them only
    // stay if every consumer CTA is already running (the usual order: the consumers are launched
    // first); otherwise they leave at once, before holding any ticket, so the consumers always fit.
    if ((int)blockIdx.x >= safe_ctas) {
        if (threadIdx.x == 0) started = ld_relaxed_u32(&ctl[523]);
        __syncthreads();
        if ((int)started < consumer_ctas) return;
    }
    int act[3], n_act = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b)
        if (wa.b[b].it1 > wa.b[b].it0) act[n_act++] = b;
    if (n_act == 0 || g_test_starve) return;
    int lo[3];
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const uint32_t *claimed = &ctl[256 + (smid & 255u)];      // set by a consumer CTA running on this SM
    // Nine ticket queues (band, channel), each in (step block, image) order.  A producer warp starts in the
    // Y : Co : Cg = a : b : c pattern (pattern = 100 a + 10 b + c, default 1 : 2 : 2) of one active band and then walks
    // all nine queues; a drawn ticket is produced as soon as the symbols it needs arrive.  (Measured alternatives
    // that changed nothing on c1: other patterns, earliest-needed-first drawing of ready tickets only, producers
    // sharing the consumers' SMs -- the step is bound by the Y -> Co -> Cg hand-over latency, not by capacity.)
    {
        bool stay = false;
        const int pa = pattern / 100, pb = pattern / 10 % 10, pc = pattern % 10, pt = pa + pb + pc;
        const int producer_id = blockIdx.x * 4 + wib;
        const int r = producer_id % (pt * n_act);
        const int slot0 = r / pt, r5 = r - pt * slot0;
        const int q0 = (slot0 == 0 ? act[0] : slot0 == 1 ? act[1] : act[2]) * 3 + (r5 < pa ? 0 : r5 < pa + pb ? 1 : 2);
        for (int c = 0; c < 9; ++c) {
            const int q = (q0 + c) % 9, band = q / 3, clr = q - 3 * band;
            const WaveBand &wb = band == 0 ? wa.b[0] : band == 1 ? wa.b[1] : wa.b[2];
            if (wb.it1 <= wb.it0) continue;
            const DecodeGeom dg = wb.dg;
            const size_t P = (size_t)dg.Hs * dg.Ws;
            const uint32_t total = (uint32_t)(wb.it1 - wb.it0) * (uint32_t)n;
            for (;;) {
                if (!stay && ld_relaxed_u32(claimed) != 0u) {                  // a consumer CTA runs on this SM (no ticket is held here):
                    uint32_t slot = 0;                                        // all but `share` producer warps leave it to the chains
                    if (lane == 0) slot = atomicAdd(&ctl[600 + (smid & 255u)], 1u);
                    if (__shfl_sync(kFull, slot, 0) >= (uint32_t)share) return;
                    stay = true;
                }
                uint32_t w = 0;
                if (lane == 0) w = atomicAdd(&ctl[514 + q], 1u);
                w = __shfl_sync(kFull, w, 0);
                if (w >= total || aborted()) break;
                const int tb = wb.it0 + (int)(w / (uint32_t)n);
                const int img = (int)(w % (uint32_t)n);
                const size_t ch = (size_t)img * 3 + clr;
                const CdfGrid g = band_grids(minmax + img * 4, clr, lo);
                uint4 *item = wb.items + (ch * dg.items_per_chain + tb) * kItemU4;
                produce_item<true>(wb.params + (size_t)img * kParamCh * P, wb.syms + (size_t)img * 3 * sym_cap, sym_cap, P, dg, clr,
                                   lo, g, 0, tb, np, item, stage[wib], lane);
                __syncwarp();
                if (lane == 0) {
                    __threadfence();
                    st_release_u32(wb.flags + ch * dg.items_per_chain + tb, 1u);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Legacy schedule (decode_impl = 1): one warp per chain does everything, Y, Co, Cg in turn.
// Kept for A/B measurements against the windowed schedules.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
decode_band_warp_kernel(const float *__restrict__ params, int16_t *__restrict__ planes,
                        const int32_t *__restrict__ minmax, DecodeGeom dg, NumericsProfile np,
                        const uint8_t *__restrict__ blob, const uint64_t *__restrict__ suboff,
                        const uint32_t *__restrict__ sublen, int total_sub) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // substream = warp
    const int img = blockIdx.y;
    if (j >= dg.S) return;
    const size_t P = (size_t)dg.Hs * dg.Ws;
    const float *pp = params + (size_t)img * kParamCh * P;
    int16_t *yb = planes + (size_t)img * 12 * P + (size_t)(3 * (dg.band + 1)) * P;
    const int32_t *mm = minmax + img * 4;
    const int lo[3] = {-127, mm[0], mm[1]};
    const int hi[3] = {128, mm[2], mm[3]};
    AcDecoder dec[3];
    CdfGrid grid[3];
#pragma unroll
    for (int clr = 0; clr < 3; ++clr) {
        const size_t e = (size_t)img * total_sub + dg.sub_first[clr] + j;
        dec[clr].init(blob + suboff[e], sublen[e]);
        grid[clr] = make_grid(lo[clr], hi[clr]);
    }
    const int n_sym = dg.n_sym;
    const bool rep_w = dg.padW && (dg.band == 0 || dg.band == 1);
    const bool rep_h = dg.padH && (dg.band == 0 || dg.band == 2);
    WarpParams cur, nxt;
    if (j < n_sym) {
        const int r = j / dg.crop_w, c = j - r * dg.crop_w;
        cur.load(pp, P, (size_t)r * dg.Ws + c, lane);
    }
    for (int i = j; i < n_sym; i += dg.S) {
        const int r = i / dg.crop_w, c = i - r * dg.crop_w;
        const size_t pidx = (size_t)r * dg.Ws + c;
        const int in = i + dg.S;
        if (in < n_sym) {          // prefetch the next position's network outputs
            const int rn = in / dg.crop_w, cn = in - rn * dg.crop_w;
            nxt.load(pp, P, (size_t)rn * dg.Ws + cn, lane);
        }
        int yv[3] = {0, 0, 0};
#pragma unroll
        for (int clr = 0; clr < 3; ++clr) {
            GmmChannel ch;
            warp_channel(cur, clr, yv[0], yv[1], np, ch);
            uint32_t c_low, c_high;
            const uint32_t target = target_fp64(dec[clr].low, dec[clr].high, dec[clr].value);
            const int sym = warp_search(ch, grid[clr], target, np, lane, c_low, c_high);
            if (in < n_sym) dec[clr].consume(c_low, c_high);
            yv[clr] = sym + lo[clr];
        }
        if (lane < 3) {
            const int16_t v = (int16_t)(lane == 0 ? yv[0] : lane == 1 ? yv[1] : yv[2]);
            int16_t *dst = yb + (size_t)lane * P + pidx;
            dst[0] = v;
            const bool last_c = rep_w && c == dg.crop_w - 1, last_r = rep_h && r == dg.crop_h - 1;
            if (last_c) dst[1] = v;
            if (last_r) dst[dg.Ws] = v;
            if (last_c && last_r) dst[dg.Ws + 1] = v;
        }
        cur = nxt;
    }
}

// ------------------------------------------------------------------------------------------
// Group schedule (interleaved-substream container, the throughput configs): a group of G lanes owns one
// chain position stream j of a band -- substream j of the Y, Co and Cg streams -- and decodes the three
// samples of position i = j + t S in turn, for t = 0, 1, ...  No windows, no items, no producer kernel: the
// lanes of the group evaluate G exact table entries q(base + l * stride) around the predicted value, the
// candidate is the last one with low + ((span * q) >> 16) <= value (torchac's search key without the
// division), and a miss walks / gallops from the window that missed.  The coder state is replicated in the
// lanes of the group.  Thousands of chains are in flight (one per ~sub_len symbols), so the instruction
// count per symbol is what matters here: G entries instead of the 32 of a window row, and the decoded
// samples go straight into the planes (with the replicate padding of the short phases).
// Neighbouring groups work on neighbouring positions (substream j codes symbols j, j + S, ...), so the
// parameter planes are read in full sectors.
// ------------------------------------------------------------------------------------------
// The 60 network outputs of a position arrive in shared memory two steps ahead of their use (cp.async, no
// registers held): the parameter planes were written by the CNN launch before and are far larger than L2, so every
// position is an HBM access of ~1 us that must not sit on the serial chain of the coder.
constexpr int kGroupStages = 3;
__device__ __forceinline__ void cp_async_f32(float *smem_dst, const float *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// GMM channel `clr` of a position from its 60 staged network outputs, with the mean coupling (LLICTI_nets.py:385-392).
__device__ __forceinline__ void staged_channel(const float *__restrict__ sp, int clr, int y0, int y1, const NumericsProfile &np,
                                               GmmChannel &c) {
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        c.sigma[m] = sp[clr * kM + m];
        c.mu[m] = sp[(3 + clr) * kM + m];
        c.w[m] = sp[(6 + clr) * kM + m];
    }
    if (clr == 1) {
        const float f0 = div255((float)y0, np);
#pragma unroll
        for (int m = 0; m < kM; ++m) c.mu[m] = __fadd_rn(c.mu[m], __fmul_rn(sp[9 * kM + m], f0));
    } else if (clr == 2) {
        const float f0 = div255((float)y0, np), f1 = div255((float)y1, np);
#pragma unroll
        for (int m = 0; m < kM; ++m)
            c.mu[m] = __fadd_rn(c.mu[m], __fadd_rn(__fmul_rn(sp[10 * kM + m], f0), __fmul_rn(sp[11 * kM + m], f1)));
    }
    gmm_prepare(c, np);
}

// One decoded symbol's interval update on the look-ahead bit source, with the closed-form renormalisation (one clz, one
// shift; next_state above) instead of AcDecoderW::consume's two-stage loop: the same registers afterwards.
__device__ __forceinline__ void consume_closed_form(AcDecoderW &d, uint32_t c_low, uint32_t c_high) {
    if (d.br.avail < 32) {                                  // BitReaderW::take's refill, so that 32 bits can be peeked
        d.br.buf |= (uint64_t)d.br.a0 << (32 - d.br.avail);
        d.br.avail += 32;
        d.br.a0 = d.br.a1;
        d.br.a1 = d.br.fetch(d.br.idx++);
    }
    const NextState ns = next_state(d.low, d.high, d.value, (uint32_t)(d.br.buf >> 32), c_low, c_high);
    d.low = ns.low; d.high = ns.high; d.value = ns.value;
    d.br.buf <<= ns.sh;
    d.br.avail -= ns.sh;
}

// Cheap stand-in for the mixture CDF, used only to GUESS where the symbol is (the decoder then verifies the guess with
// exact table entries, so a wrong guess costs time, never correctness): the normal CDF as
// 1 / (1 + exp(-2 c (z + 0.044715 z^3))), c = sqrt(2 / pi) -- absolute error < 3e-4 -- with MUFU.EX2 and MUFU.RCP: ten
// instructions per mixture against ~70 for the exact erfcf term.  On the synthetic images it puts the symbol exactly in
// 99.7 % of the cases and within one symbol in all of them (tools: see DESIGN.md), where a window centred on the
// mixture mean holds it in 70-96 %.  Returns the approximate table entry q~(k) as a float.
__device__ __forceinline__ float approx_q(const GmmChannel &c, const CdfGrid &g, int k) {
    if (k <= 0) return 0.f;
    if (k >= g.Lp - 1) return 65536.f;
    const float p = ((float)(g.min_val + k) - 0.5f) * (1.0f / 255.0f);
    float acc = 0.f;
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        const float z = (p - c.mu[m]) * c.rinv[m];
        const float y = z * fmaf(z * z, 0.044715f, 1.0f);
        float e, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * -2.3022082f));        // exp(-2 c y) = 2^(-2 c log2(e) y)
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
        acc = fmaf(c.w[m], r, acc);
    }
    return fmaf(acc, g.scale, (float)k);
}

// First probes of the approximate search, relative to the index of the mixture mean: geometric spacing, so that one
// round brackets the (usual) near symbols tightly and the (rare) far ones at all.
template <int G> __device__ __forceinline__ int first_probe_offset(int l);
template <> __device__ __forceinline__ int first_probe_offset<2>(int l) { return l == 0 ? -2 : 3; }
template <> __device__ __forceinline__ int first_probe_offset<4>(int l) { return l == 0 ? -12 : l == 1 ? -2 : l == 2 ? 3 : 13; }
template <> __device__ __forceinline__ int first_probe_offset<8>(int l) {
    return l == 0 ? -40 : l == 1 ? -13 : l == 2 ? -4 : l == 3 ? -1 : l == 4 ? 2 : l == 5 ? 5 : l == 6 ? 14 : 41;
}
template <> __device__ __forceinline__ int first_probe_offset<16>(int l) { return 3 * (l - 8) + 1; }

// The same, shared out over the lanes of a group of G >= 8: lane m < 5 prepares mixture m (clamps, coupling, the
// weight's division, the refined reciprocal of the spread), then the 20 prepared values are exchanged by shuffles --
// a third of the instructions of every lane preparing all five mixtures.  Bit-identical to gmm_prepare (same operations
// on the same operands; the weight normaliser is summed by every lane in the profile's order).
template <int G>
__device__ __forceinline__ void staged_channel_shared(const float *__restrict__ sp, int clr, int y0, int y1, const NumericsProfile &np,
                                                      int sub, GmmChannel &c) {
    static_assert(G >= 8, "needs five lanes per group");
    const int m = min(sub, kM - 1);
    const float sb = (float)(0.11 / 255.0);
    float wall[kM];
#pragma unroll
    for (int k = 0; k < kM; ++k) wall[k] = fmaxf(sp[(6 + clr) * kM + k], 1e-6f);
    const float den = __fadd_rn(sum5(wall, np), 1e-9f);
    float sg = fmaxf(sp[clr * kM + m], sb), mu = sp[(3 + clr) * kM + m];
    if (clr == 1) {
        mu = __fadd_rn(mu, __fmul_rn(sp[9 * kM + m], div255((float)y0, np)));
    } else if (clr == 2) {
        mu = __fadd_rn(mu, __fadd_rn(__fmul_rn(sp[10 * kM + m], div255((float)y0, np)), __fmul_rn(sp[11 * kM + m], div255((float)y1, np))));
    }
    const float w = __fdiv_rn(fmaxf(sp[(6 + clr) * kM + m], 1e-6f), den);
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(sg));
    const float rinv = __fmaf_rn(r0, __fmaf_rn(-sg, r0, 1.0f), r0);
    const bool fast = (fabsf(mu) <= 1024.0f) & (sg >= 9.5367431640625e-07f) & (sg <= 1048576.0f);
#pragma unroll
    for (int k = 0; k < kM; ++k) {
        c.sigma[k] = __shfl_sync(kFull, sg, k, G);
        c.mu[k] = __shfl_sync(kFull, mu, k, G);
        c.w[k] = __shfl_sync(kFull, w, k, G);
        c.rinv[k] = __shfl_sync(kFull, rinv, k, G);
    }
    const unsigned slow = __ballot_sync(kFull, !fast && sub < kM);        // a group's lanes 0..4 vote for their mixtures
    c.fast = ((slow >> ((threadIdx.x & 31) - sub)) & ((1u << G) - 1u)) == 0u;
}

template <int G, bool kLocate>
__global__ void __launch_bounds__(128)
decode_band_group_kernel(const float *__restrict__ params, int16_t *__restrict__ planes, const int32_t *__restrict__ minmax,
                         DecodeGeom dg, NumericsProfile np, const uint8_t *__restrict__ blob,
                         const uint64_t *__restrict__ suboff, const uint32_t *__restrict__ sublen, int total_sub, int n) {
    static_assert(G == 2 || G == 4 || G == 8 || G == 16, "group size");
    __shared__ float stage[kGroupStages][128 / G][kParamCh];
    const int lane = threadIdx.x & 31, sub = lane & (G - 1);
    const int gshift = lane - sub;
    constexpr uint32_t gm = (1u << G) - 1u;
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int gib = threadIdx.x / G;                 // group within the block
    const bool valid = gid < (long long)n * dg.S;
    // whole warps without work leave; otherwise every lane stays for the warp-wide ballots and shuffles
    if (!__any_sync(kFull, valid)) return;
    const int img = valid ? (int)(gid / dg.S) : 0, j = valid ? (int)(gid - (long long)img * dg.S) : 0;
    const size_t P = (size_t)dg.Hs * dg.Ws;
    const float *pp = params + (size_t)img * kParamCh * P;
    int16_t *yb = planes + (size_t)img * 12 * P + (size_t)(3 * (dg.band + 1)) * P;
    const int32_t *mm = minmax + img * 4;
    const int mn_co = mm[0], mn_cg = mm[1], mx_co = mm[2], mx_cg = mm[3];
    const CdfGrid gY = make_grid(-127, 128), gCo = make_grid(mn_co, mx_co), gCg = make_grid(mn_cg, mx_cg);
    // The decoder of the channel being decoded is always `d`: the three are rotated after every symbol, so the body
    // below exists once in the code (a tenth of the unrolled form's instruction footprint).
    AcDecoderW d, d_next, d_last;
    {
        const size_t e0 = (size_t)img * total_sub + j;
        d.init(blob + suboff[e0 + dg.sub_first[0]], valid ? sublen[e0 + dg.sub_first[0]] : 0u);
        d_next.init(blob + suboff[e0 + dg.sub_first[1]], valid ? sublen[e0 + dg.sub_first[1]] : 0u);
        d_last.init(blob + suboff[e0 + dg.sub_first[2]], valid ? sublen[e0 + dg.sub_first[2]] : 0u);
    }
    const bool rep_w = dg.padW && (dg.band == 0 || dg.band == 1);
    const bool rep_h = dg.padH && (dg.band == 0 || dg.band == 2);
    const int n_steps = valid ? (dg.n_sym - j + dg.S - 1) / dg.S : 0;
    const int max_steps = dg.max_steps;          // warp-uniform trip count
    auto prefetch = [&](int t) {                 // the group's lanes share the 60 loads of step t's position
        if (t < n_steps) {
            const int i = j + t * dg.S;
            const int r = i / dg.crop_w, c = i - r * dg.crop_w;
            const float *src = pp + (size_t)r * dg.Ws + c;
            float *dst = stage[t % kGroupStages][gib];
            for (int ch = sub; ch < kParamCh; ch += G) cp_async_f32(dst + ch, src + (size_t)ch * P);
        }
        cp_async_commit();
    };
    prefetch(0);
    prefetch(1);
    unsigned long long rounds = 0;
    int t = 0, clr = 0, y0 = 0, y1 = 0, r = 0, c = 0;
    bool live = false;
    const float *sp = stage[0][gib];
#pragma unroll 1
    for (;;) {
        if (clr == 0) {
            if (t >= max_steps) break;
            cp_async_wait<1>();                      // step t has landed (step t+1 may be in flight)
            __syncwarp();                            // ... for every lane of the group; and step t-1's reads are done
            prefetch(t + 2);
            live = t < n_steps;
            const int i = j + t * dg.S;
            r = live ? i / dg.crop_w : 0;
            c = live ? i - r * dg.crop_w : 0;
            sp = stage[t % kGroupStages][gib];
        }
        const int lo_c = clr == 0 ? -127 : clr == 1 ? mn_co : mn_cg;
        CdfGrid g;                                           // (the grids' end points cost two fp64 divisions: made once, above)
        g.min_val = lo_c;
        g.Lp = clr == 0 ? gY.Lp : clr == 1 ? gCo.Lp : gCg.Lp;
        g.p_first = clr == 0 ? gY.p_first : clr == 1 ? gCo.p_first : gCg.p_first;
        g.p_last = clr == 0 ? gY.p_last : clr == 1 ? gCo.p_last : gCg.p_last;
        g.scale = clr == 0 ? gY.scale : clr == 1 ? gCo.scale : gCg.scale;
        const int last = g.Lp - 1;
        GmmChannel ch;
        if (G >= 8) {                                        // (warp-wide shuffles inside: every lane, live or not)
            if constexpr (G >= 8) staged_channel_shared<G>(sp, clr, y0, y1, np, sub, ch);
        } else if (live) {
            staged_channel(sp, clr, y0, y1, np, ch);
        }
        if (!live) {
#pragma unroll
            for (int m = 0; m < kM; ++m) { ch.sigma[m] = 1.f; ch.mu[m] = 0.f; ch.w[m] = 0.2f; ch.rinv[m] = 1.f; }
            ch.fast = 1;
        }
        float mean = 0.f;
#pragma unroll
        for (int m = 0; m < kM; ++m) mean = fmaf(ch.w[m], ch.mu[m], mean);
        const int kc = __float2int_rn(mean * 255.0f) - g.min_val;
        const uint32_t low = d.low, sm1 = d.high - d.low, span = sm1 + 1u, value32 = d.value;
        // ---- guess: largest k with q~(k) <= target, by a (G+1)-ary search of the group's lanes over the approximate table ----
        int guess = kc;
        if (kLocate) {
            // torchac's search key ((value - low + 1) 2^16 - 1) / span, in fp32 (7 digits: a guess is all it feeds)
            const float tf = __fdividef(((float)(d.value - d.low) + 1.0f) * 65536.0f, (float)sm1 + 1.0f);
            int a = 0, bnd = last;                           // the answer lies in [a, bnd)
            bool located = !live;
            {
                const int k = kc + first_probe_offset<G>(sub);
                const bool le = k <= 0 || (k < last && approx_q(ch, g, k) <= tf);
                const int cnt = __popc((__ballot_sync(kFull, le) >> gshift) & gm);
                const int k_lo = __shfl_sync(kFull, k, max(cnt - 1, 0), G), k_hi = __shfl_sync(kFull, k, min(cnt, G - 1), G);
                if (cnt > 0) a = min(max(k_lo, 0), last - 1);
                if (cnt < G) bnd = max(min(k_hi, last), a + 1);
            }
#pragma unroll 1
            for (;;) {
                located = located || bnd - a <= 1;
                if (__all_sync(kFull, located)) break;
                const int step = max((bnd - a + G) / (G + 1), 1);
                const int k = a + (sub + 1) * step;
                const bool le = !located && k < bnd && approx_q(ch, g, k) <= tf;
                const int cnt = __popc((__ballot_sync(kFull, le) >> gshift) & gm);
                if (!located) {
                    bnd = min(bnd, a + (cnt + 1) * step);
                    a += cnt * step;
                }
            }
            guess = a;
        }
        // ---- verify with exact entries.  q(s_lo) <= target is known (or s_lo = 0), q(s_hi) > target is known (s_hi = last: 2^16)
        int s_lo = 0, s_hi = last, base = min(max(guess - (G / 2 - 1), 0), max(last - (G - 1), 0)), stride = 1, round = 0;
        bool done = !live, miss_down = false;
        uint32_t c_low = 0, c_high = 0x10000u;
        int sym = 0;
#pragma unroll 1
        for (;;) {
            const int k = base + sub * stride;
            uint32_t q = 0x10000u;
            if (!done && k < last) q = cdf_q(ch, g, k, np);
            // candidate test low + ((span * q) >> 16) <= value in 32 bits: the high word of span * (q << 16); the full range
            // (span = 2^32, seen as 0: a stream's first symbols) gives q << 16 itself, and q = 2^16 (the alphabet's end, or a
            // probe beyond it) gives high + 1 > value
            const uint32_t q16 = q << 16;
            const uint32_t nl = low + (span == 0u ? q16 : __umulhi(span, q16));
            const unsigned le = __ballot_sync(kFull, q != 0x10000u && nl <= value32);
            const int cnt = __popc((le >> gshift) & gm);
            const int i_lo = max(cnt - 1, 0), i_hi = min(i_lo + 1, G - 1);
            const uint32_t q_lo = __shfl_sync(kFull, q, i_lo, G), q_hi = __shfl_sync(kFull, q, i_hi, G);
            if (!done) {
                ++round;
                if (cnt == 0) {
                    if (base == 0 && stride == 1) {        // below q(0): torchac's search returns symbol 0
                        sym = 0; c_low = q_lo; c_high = q_hi; done = true;
                    } else {
                        s_hi = max(base, 1);
                        miss_down = true;
                    }
                } else if (cnt == G) {
                    s_lo = base + (G - 1) * stride;
                    miss_down = false;
                } else {
                    s_lo = base + (cnt - 1) * stride;
                    s_hi = min(base + cnt * stride, last);
                    miss_down = false;
                    if (stride == 1) { sym = s_lo; c_low = q_lo; c_high = q_hi; done = true; }
                }
                if (!done) {
                    // next probes: the entries next to the window that missed, then four times as far, then whatever is
                    // left of [s_lo, s_hi] in G equal steps (probe 0 repeats s_lo, the others lie strictly inside: every
                    // round narrows the interval, also for G = 2)
                    const int full = max((s_hi - s_lo + G - 1) / G, 1);
                    stride = min(full, round == 1 ? 1 : round == 2 ? 4 : full);
                    base = miss_down ? max(s_hi - (G - 1) * stride, s_lo) : s_lo;
                }
            }
            if (__all_sync(kFull, done)) break;
        }
        if (live) {
            if (t + 1 < n_steps) consume_closed_form(d, c_low, c_high);      // torchac does not update after the last symbol
            rounds += (unsigned long long)(round - 1);
        }
        const int yv = sym + lo_c;
        if (live && sub == 0) {                                  // the sample goes straight into the planes, with the replicate padding
            const int16_t v = (int16_t)yv;
            int16_t *dst = yb + (size_t)clr * P + (size_t)r * dg.Ws + c;
            dst[0] = v;
            const bool last_c = rep_w && c == dg.crop_w - 1, last_r = rep_h && r == dg.crop_h - 1;
            if (last_c) dst[1] = v;
            if (last_r) dst[dg.Ws] = v;
            if (last_c && last_r) dst[dg.Ws + 1] = v;
        }
        // rotate: the next channel's decoder becomes `d`
        { const AcDecoderW tmp = d; d = d_next; d_next = d_last; d_last = tmp; }
        if (clr == 0) { y0 = yv; clr = 1; }
        else if (clr == 1) { y1 = yv; clr = 2; }
        else { clr = 0; ++t; }
    }
    cp_async_wait<0>();
    if (sub == 0 && rounds) atomicAdd(&g_decode_stats[0], rounds);       // search rounds beyond the first (window misses)
}

// ------------------------------------------------------------------------------------------
// Lane schedule (default for the interleaved-substream container): ONE lane per coded substream.
//
// The group schedule above spends G lanes on one symbol -- every lane prepares the same mixture, runs the same
// approximate search and evaluates an exact table entry of which at most two are used: ~4000 thread-instructions per
// symbol at G = 4, and the launch is issue-bound (62 % issue utilisation at 4 warps per scheduler).  Here a lane does the
// minimum a decoded symbol needs: prepare ITS channel's mixture, find the symbol by a safeguarded Newton iteration on
// the cheap logistic stand-in of the CDF (scalar: no ballots or shuffles), then evaluate the two exact entries q(s),
// q(s + 1) that prove the symbol and are its interval (the same two entries the encoder evaluates).
// The parallelism the group's lanes gave comes back from the colour channels: substream j of the Y, Co and Cg streams
// are three lanes of one warp running one and two steps apart -- lane Y decodes position t while lane Co decodes t - 1
// with the Y sample it was handed by a shuffle and lane Cg decodes t - 2 -- so 3 x n x S lanes are in flight instead of
// n x S groups.  A warp holds 10 chains: lanes 0-9 Y, 10-19 Co, 20-29 Cg (neighbouring lanes = neighbouring positions,
// so parameter loads and sample stores touch whole sectors); lanes 30, 31 idle.
// A wrong guess costs time, never correctness: the exact pair test decides, and a miss widens the search with exact
// entries (neighbours first, then four apart, then thirds of what is left).
// ------------------------------------------------------------------------------------------
constexpr int kLaneChains = 10;
constexpr int kLaneStages = 3;
constexpr int kLaneParams = 5 * kM;     // 15 mixture parameters + up to 10 coupling coefficients
#ifndef LLICTI_LANE_MARGIN
#define LLICTI_LANE_MARGIN 0.2f
#endif
constexpr float kLaneMargin = LLICTI_LANE_MARGIN;

// Stand-in table entry q~(k), its slope and its curvature per index step, 1 <= k <= Lp - 2.  The normal CDF through
// Abramowitz & Stegun 7.1.26, erfc(x) = (a1 t + ... + a5 t^5) exp(-x^2), t = 1 / (1 + p x), |error| < 1.5e-7 -- MUFU.RCP +
// MUFU.EX2 + five FMAs per mixture, a third of libdevice's erfcf -- and the same exponential is the density (the slope) and,
// times -z / sigma, its derivative (the curvature).  Still a GUESS: the exact entries decide.
__device__ __forceinline__ void approx_q_slope(const GmmChannel &c, const CdfGrid &g, int k, float &q, float &dq, float &ddq) {
    const float p = ((float)(g.min_val + k) - 0.5f) * (1.0f / 255.0f);
    float acc = 0.f, dacc = 0.f, cacc = 0.f;
#pragma unroll
    for (int m = 0; m < kM; ++m) {
        const float z = (p - c.mu[m]) * c.rinv[m];
        const float x = fabsf(z) * 0.70710678f;
        float e, t;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -1.4426950f));              // exp(-x^2) = exp(-z^2 / 2)
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, x, 1.0f)));
        const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
        const float h = 0.5f * poly * e;                                                     // 0.5 erfc(|z| / sqrt 2)
        acc = fmaf(c.w[m], z >= 0.f ? 1.0f - h : h, acc);
        const float wd = c.w[m] * c.rinv[m] * e;                                             // density / 0.39894228
        dacc += wd;
        cacc = fmaf(wd * c.rinv[m], -z, cacc);                                               // its derivative: -z / sigma times it
    }
    q = fmaf(acc, g.scale, (float)k);
    dq = fmaf(dacc, g.scale * (0.39894228f / 255.0f), 1.0f);
    ddq = cacc * (g.scale * (0.39894228f / (255.0f * 255.0f)));
}

// kOcc = CTAs per SM the register allocation aims at.  A band of the finest scale of a c2 / c3 batch is ~ 24 warps of chains
// per SM: with kOcc = 6 (80 registers, 4 bytes of spill) ALL of them are resident in one wave -- six warps per scheduler hide
// the dependent-issue latency that bounds the kernel -- where 4 CTAs per SM at 118 registers ran a full wave and a half-empty
// one.  The staging buffer holds the 30 working lanes of a warp only (36,000 B: six CTAs fit the SM's shared memory).
constexpr int kLaneSlots = 4 * 3 * kLaneChains;   // working lanes of a CTA
template <int kOcc>
__global__ void __launch_bounds__(128, kOcc)
decode_band_lane_kernel(const float *__restrict__ params, int16_t *__restrict__ planes, const int32_t *__restrict__ minmax,
                        DecodeGeom dg, NumericsProfile np, const uint8_t *__restrict__ blob,
                        const uint64_t *__restrict__ suboff, const uint32_t *__restrict__ sublen, int total_sub, int n, int guess_skew) {
    __shared__ float stage[kLaneStages][kLaneParams][kLaneSlots];
    const int lane = threadIdx.x & 31;
    const int slot = min((int)(threadIdx.x >> 5) * 3 * kLaneChains + lane, kLaneSlots - 1);   // (idle lanes 30, 31 never stage)
    const int clr = lane / kLaneChains;                       // 3: idle lane
    const long long wg = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long gid = wg * kLaneChains + (lane - clr * kLaneChains);
    const bool valid = clr < 3 && gid < (long long)n * dg.S;
    if (!__any_sync(kFull, valid)) return;
    const int img = valid ? (int)(gid / dg.S) : 0, j = valid ? (int)(gid - (long long)img * dg.S) : 0;
    const size_t P = (size_t)dg.Hs * dg.Ws;
    const int cl = min(clr, 2);
    const float *pp = params + (size_t)img * kParamCh * P;
    int16_t *yb = planes + (size_t)img * 12 * P + (size_t)(3 * (dg.band + 1) + cl) * P;
    const int32_t *mm = minmax + img * 4;
    const CdfGrid g = cl == 0 ? make_grid(-127, 128) : cl == 1 ? make_grid(mm[0], mm[2]) : make_grid(mm[1], mm[3]);
    const int last = g.Lp - 1;
    AcDecoderW d;
    {
        const size_t e = (size_t)img * total_sub + j + dg.sub_first[cl];
        d.init(blob + suboff[e], valid ? sublen[e] : 0u);
    }
    const bool rep_w = dg.padW && (dg.band == 0 || dg.band == 1);
    const bool rep_h = dg.padH && (dg.band == 0 || dg.band == 2);
    const int n_steps = valid ? (dg.n_sym - j + dg.S - 1) / dg.S : 0;
    // network-output planes that feed this lane's channel: sigma, mu, weight (five mixtures each), then its coupling
    // coefficients (Co: one set, Cg: two); staged parameter k sits at stage[.][k][thread]
    const float *p_sigma = pp + (size_t)(cl * kM) * P, *p_mu = pp + (size_t)((3 + cl) * kM) * P, *p_w = pp + (size_t)((6 + cl) * kM) * P;
    const float *p_cpl = pp + (size_t)((cl == 1 ? 9 : 10) * kM) * P;
    // position of step t: i = j + t S -> (row, column) of the cropped band, advanced by (S / crop_w, S % crop_w) per step
    // (two counters: the prefetch runs two steps ahead of the decode)
    const int step_r = dg.S / dg.crop_w, step_c = dg.S - step_r * dg.crop_w;
    int pf_r = j / dg.crop_w, pf_c = j - pf_r * dg.crop_w;
    int dec_r = pf_r, dec_c = pf_c;
    auto prefetch = [&](int tau) {                            // parameters consumed in iteration tau (this lane's step tau - clr)
        const int t = tau - cl;
        if (t >= 0 && t < n_steps) {
            const size_t off = (size_t)pf_r * dg.Ws + pf_c;
            pf_c += step_c; pf_r += step_r;
            if (pf_c >= dg.crop_w) { pf_c -= dg.crop_w; ++pf_r; }
            float *dst = &stage[tau % kLaneStages][0][slot];
#pragma unroll
            for (int m = 0; m < kM; ++m) {
                cp_async_f32(dst + m * kLaneSlots, p_sigma + off + (size_t)m * P);
                cp_async_f32(dst + (kM + m) * kLaneSlots, p_mu + off + (size_t)m * P);
                cp_async_f32(dst + (2 * kM + m) * kLaneSlots, p_w + off + (size_t)m * P);
            }
            if (cl >= 1) {
#pragma unroll
                for (int m = 0; m < kM; ++m) cp_async_f32(dst + (3 * kM + m) * kLaneSlots, p_cpl + off + (size_t)m * P);
            }
            if (cl == 2) {
#pragma unroll
                for (int m = 0; m < kM; ++m) cp_async_f32(dst + (4 * kM + m) * kLaneSlots, p_cpl + off + (size_t)(kM + m) * P);
            }
        }
        cp_async_commit();
    };
    prefetch(0);
    prefetch(1);
    unsigned long long extra_rounds = 0, newton_its = 0;
    int in0 = 0, in1 = 0;                                     // decoded Y (and Co) sample of this lane's position
    const int iters = dg.max_steps + 2;                       // warp-uniform
#pragma unroll 1
    for (int tau = 0; tau < iters; ++tau) {
        cp_async_wait<1>();                                   // this iteration's parameters have landed (own copies only: no barrier)
        prefetch(tau + 2);
        const int t = tau - cl;
        const bool live = t >= 0 && t < n_steps;
        int yv = 0;
        if (live) {
            const float *sp = &stage[tau % kLaneStages][0][slot];
            GmmChannel ch;
#pragma unroll
            for (int m = 0; m < kM; ++m) {
                ch.sigma[m] = sp[m * kLaneSlots];
                ch.mu[m] = sp[(kM + m) * kLaneSlots];
                ch.w[m] = sp[(2 * kM + m) * kLaneSlots];
            }
            if (cl == 1) {                                    // mean coupling (LLICTI_nets.py:385-392), as staged_channel
                const float f0 = div255((float)in0, np);
#pragma unroll
                for (int m = 0; m < kM; ++m) ch.mu[m] = __fadd_rn(ch.mu[m], __fmul_rn(sp[(3 * kM + m) * kLaneSlots], f0));
            } else if (cl == 2) {
                const float f0 = div255((float)in0, np), f1 = div255((float)in1, np);
#pragma unroll
                for (int m = 0; m < kM; ++m)
                    ch.mu[m] = __fadd_rn(ch.mu[m], __fadd_rn(__fmul_rn(sp[(3 * kM + m) * kLaneSlots], f0), __fmul_rn(sp[(4 * kM + m) * kLaneSlots], f1)));
            }
            gmm_prepare(ch, np);
            const uint32_t low = d.low, sm1 = d.high - d.low, span = sm1 + 1u, value32 = d.value;
            // ---- guess: floor of the root of q~(x) = target, by Newton steps kept inside a bracket [a, b) -------------
            int guess = 0;
            if (last >= 2) {
                // torchac's search key ((value - low + 1) 2^16 - 1) / span, in fp32 (a guess is all it feeds)
                const float tf = __fdividef(((float)(value32 - low) + 1.0f) * 65536.0f, (float)sm1 + 1.0f);
                // first probe: the target's quantile of the Gaussian with the mixture's mean and variance, through the same
                // logistic stand-in -- y = logit(u) / 2c, z from z + 0.044715 z^3 = y by one Newton step off z0 = y / (1 + 0.044715 y^2)
                // (the mixture mean alone is the right start only for targets near one half: one iteration more on average)
                float mean = 0.f, m2 = 0.f;
#pragma unroll
                for (int m = 0; m < kM; ++m) {
                    mean = fmaf(ch.w[m], ch.mu[m], mean);
                    m2 = fmaf(ch.w[m], fmaf(ch.sigma[m], ch.sigma[m], ch.mu[m] * ch.mu[m]), m2);
                }
                const float sd = sqrtf(fmaxf(fmaf(-mean, mean, m2), 1e-12f));
                float smin = 1.0f;                                                               // over the components that carry weight
#pragma unroll
                for (int m = 0; m < kM; ++m) smin = fminf(smin, ch.w[m] >= 0.004f ? ch.sigma[m] : 1.0f);
                const bool smooth = smin >= 0.75f / 255.0f;                                       // every such component at least 3/4 of a level wide
                const float u = fminf(fmaxf(tf * (1.0f / 65536.0f), 1e-6f), 1.0f - 1e-6f);
                const float y = (__log2f(u) - __log2f(1.0f - u)) * 0.43436f;                     // ln 2 / (2 c)
                const float z0 = __fdividef(y, fmaf(0.044715f * y, y, 1.0f));
                const float zq = z0 - __fdividef(fmaf(0.044715f * z0 * z0, z0, z0) - y, fmaf(0.134145f * z0, z0, 1.0f));
                int a = 0, b = last, k = __float2int_rd(fmaf(zq, sd, mean) * 255.0f + 0.5f) - g.min_val;
#pragma unroll 1
                for (int it = 0;; ++it) {
                    if (b - a <= 1) { guess = a; break; }
                    k = min(max(k, a + 1), b - 1);
                    float q, dq, ddq;
                    approx_q_slope(ch, g, k, q, dq, ddq);
                    // Halley step: the Newton step corrected by the curvature (res = dq dx + ddq dx^2 / 2).  In the tails of a
                    // peaky mixture the CDF bends so much within one symbol that the plain Newton step -- accepted when it
                    // lands inside the neighbouring symbol -- was one symbol short in 1.5 % of the cases with trained weights.
                    const float res = tf - q, dxn = __fdividef(res, dq);
                    const float dx = __fdividef(res, fmaxf(fmaf(0.5f * ddq, dxn, dq), 0.3f * dq));
                    ++newton_its;
                    // The estimated root k + dx decides as long as it is not within kLaneMargin of the bound the bracket has not
                    // confirmed; otherwise the neighbouring entry is evaluated too.  Only for smooth mixtures: with a component
                    // near the spread clamp the CDF is a staircase, a local slope says nothing about the next entry, and the
                    // bracket alone decides.
                    if (res >= 0.f) {
                        a = k;
                        if (b == k + 1 || (smooth && dx < 1.0f - kLaneMargin)) { guess = k; break; }          // root in [k, k + 1)
                    } else {
                        b = k;
                        if (a == k - 1 || (smooth && dx >= -1.0f + kLaneMargin)) { guess = k - 1; break; }    // root in (k - 1, k)
                    }
                    int kn = k + (int)floorf(fminf(fmaxf(res >= 0.f ? dx + kLaneMargin : dx, -70000.f), 70000.f));
                    // Peaky mixtures (trained weights: one component of 3/4 of the weight a level wide, narrow ones at the spread
                    // clamp, a broad one of a few percent) bend so hard that a step taken from the flank overshoots into the
                    // flat tail by a hundred symbols, and bisection then needs seven evaluations to come back: the step is
                    // limited to 2, 4, 8, ... symbols, and moves at least one (tools/search_sim.py: the slowest of a warp's 30
                    // lanes needs 3.9 evaluations instead of 7.1).
                    if (!smooth) {
                        const int lim = 2 << min(it, 8);
                        kn = min(max(kn, k - lim), k + lim);
                        if (kn == k) kn = res >= 0.f ? k + 1 : k - 1;
                    }
                    k = (kn <= a || kn >= b || it >= 8) ? (a + b) >> 1 : kn;
                }
            }
            if (guess_skew) guess = min(max(guess + (int)((unsigned)(tau * 7 + lane) % (unsigned)(2 * guess_skew + 1)) - guess_skew, 0), last - 1);   // tests: wrong guesses
            // ---- prove: exact entries.  Invariant: entry s_lo passes the candidate test (or s_lo = 0: torchac's search
            // returns symbol 0 below q(0)), entry s_hi fails it (s_hi = last: 2^16); done when they are neighbours and
            // q(s_lo) has been evaluated.
            int s_lo = 0, s_hi = last, ka = guess, kb = guess + 1, round = 0, dir = 0;
            uint32_t q_lo = 0u, q_hi = 0x10000u;
            bool lo_known = false, single = false;     // single: this round evaluates entry ka only
#pragma unroll 1
            for (;;) {
                const uint32_t qa = cdf_q(ch, g, ka, np);
                const uint32_t qb = !single && kb < last ? cdf_q(ch, g, kb, np) : 0x10000u;
                // candidate test low + ((span * q) >> 16) <= value in 32 bits: the high word of span * (q << 16); the full
                // range (span = 2^32, seen as 0) gives q << 16 itself
                const uint32_t qa16 = qa << 16, qb16 = qb << 16;
                const bool va = ka == 0 || low + (span == 0u ? qa16 : __umulhi(span, qa16)) <= value32;
                const bool vb = !single && kb < last && low + (span == 0u ? qb16 : __umulhi(span, qb16)) <= value32;
                if (!va) { s_hi = ka; q_hi = qa; dir = -1; }
                else if (single) { s_lo = ka; q_lo = qa; lo_known = true; }
                else if (!vb) { s_lo = ka; q_lo = qa; lo_known = true; s_hi = kb; q_hi = qb; }
                else { s_lo = kb; q_lo = qb; lo_known = true; dir = 1; }
                ++round;
                if (s_hi - s_lo == 1 && lo_known) break;
                // A guess that was one off -- the usual miss -- needs ONE more entry (the other bound is already exact): the
                // second round evaluates just the neighbour; later rounds probe pairs four apart, then thirds of what is left.
                single = round == 1;
                if (single) {
                    ka = dir < 0 ? max(s_hi - 1, 0) : s_lo + 1;
                    continue;
                }
                const int lo_min = lo_known ? s_lo + 1 : s_lo, hi_max = s_hi - 1;     // entries still worth evaluating (lo_min <= hi_max)
                const int step = round == 2 ? 4 : max((s_hi - s_lo) / 3, 1);
                if (dir < 0) { kb = max(s_hi - step, lo_min); ka = max(kb - step, lo_min); }
                else { ka = min(s_lo + step, hi_max); kb = min(ka + step, hi_max); }
                if (ka == kb) {
                    if (kb < hi_max) ++kb;
                    else if (ka > lo_min) --ka;
                    else kb = ka + 1;                                               // = s_hi: re-evaluated (or the alphabet's end)
                }
            }
            extra_rounds += (unsigned long long)(round - 1);
            if (t + 1 < n_steps) consume_closed_form(d, q_lo, q_hi);                // torchac does not update after the last symbol
            yv = s_lo + g.min_val;
            {                                                                       // the sample goes straight into the planes, with the replicate padding
                const int r = dec_r, c = dec_c;
                dec_c += step_c; dec_r += step_r;
                if (dec_c >= dg.crop_w) { dec_c -= dg.crop_w; ++dec_r; }
                const int16_t v = (int16_t)yv;
                int16_t *dst = yb + (size_t)r * dg.Ws + c;
                dst[0] = v;
                const bool last_c = rep_w && c == dg.crop_w - 1, last_r = rep_h && r == dg.crop_h - 1;
                if (last_c) dst[1] = v;
                if (last_r) dst[dg.Ws] = v;
                if (last_c && last_r) dst[dg.Ws + 1] = v;
            }
        }
        // hand the samples on: Co receives Y's, Cg receives Co's pair (the Y sample Co was working with and Co's own)
        const int from_dec = __shfl_up_sync(kFull, yv, kLaneChains), from_in0 = __shfl_up_sync(kFull, in0, kLaneChains);
        if (cl == 1) in0 = from_dec;
        else if (cl == 2) { in0 = from_in0; in1 = from_dec; }
    }
    cp_async_wait<0>();
    if (extra_rounds) atomicAdd(&g_decode_stats[0], extra_rounds);       // exact rounds beyond the first (wrong guesses)
    if (newton_its) atomicAdd(&g_decode_stats[1], newton_its);          // stand-in evaluations of the guess
}

// ------------------------------------------------------------------------------------------
// Launcher
// ------------------------------------------------------------------------------------------
static int stream_index(const Plan &p, int scale, int band, int clr) {
    return (p.g.num_scales - 1 - scale) * 9 + 3 * band + clr;
}

static DecodeGeom make_decode_geom(const Plan &p, int scale, int band) {
    DecodeGeom dg;
    const StreamDesc &d0 = p.sd[stream_index(p, scale, band, 0)];
    dg.Hs = d0.Hs; dg.Ws = d0.Ws; dg.crop_h = d0.crop_h; dg.crop_w = d0.crop_w; dg.band = band;
    dg.padH = p.g.padH[scale]; dg.padW = p.g.padW[scale];
    dg.n_sym = d0.n_sym;
    dg.S = d0.S;
    dg.max_steps = (d0.n_sym + d0.S - 1) / d0.S;
    dg.items_per_chain = (dg.max_steps + 31) / 32;
    for (int c = 0; c < 3; ++c) dg.sub_first[c] = p.sd[stream_index(p, scale, band, c)].sub_first;
    return dg;
}

// Turns a raised g_abort into the context's status flag (and clears it); last launch of every decode that used a
// schedule with kernel-to-kernel hand-overs.
int launch_abort_check(llicti_ctx *ctx, cudaStream_t st) {
    abort_to_status_kernel<<<1, 1, 0, st>>>(ctx->d_status);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

// Test hooks, read once per context creation: LLICTI_TEST_POLL_LIMIT shortens the bounded waits, LLICTI_TEST_STARVE
// makes every producer leave at once (what a producer kernel that cannot become resident looks like to the consumers).
int apply_decode_test_knobs() {
    const char *pl = getenv("LLICTI_TEST_POLL_LIMIT"), *sv = getenv("LLICTI_TEST_STARVE");
    const uint32_t polls = pl && *pl ? (uint32_t)atoll(pl) : (1u << 25);
    const int starve = sv && *sv ? atoi(sv) : 0;
    LLICTI_CUDA(cudaMemcpyToSymbol(g_max_polls, &polls, sizeof(polls)));
    LLICTI_CUDA(cudaMemcpyToSymbol(g_test_starve, &starve, sizeof(starve)));
    return LLICTI_OK;
}

size_t decode_item_bytes() { return (size_t)kItemU4 * sizeof(uint4); }

// Window items one image needs for its largest band.
int64_t decode_flag_words(int64_t items_cap) { return items_cap + kCtlWords; }

int64_t decode_items_per_image(const Plan &p) {
    int64_t m = 0;
    for (int s = 0; s < p.g.num_scales; ++s)
        for (int b = 0; b < 3; ++b) {
            const DecodeGeom dg = make_decode_geom(p, s, b);
            m = std::max<int64_t>(m, 3ll * dg.S * dg.items_per_chain);
        }
    return m;
}

static int env_int(const char *name, int dflt);
static int group_lanes() { return std::max(env_int("LLICTI_GROUP_LANES", 4), 2); }
static long long group_min_warps() { return env_int("LLICTI_GROUP_MIN_WARPS", 592); }      // one warp per scheduler of a B200

// Does band (scale, b) of a batch of n images take the group schedule?  (Substream container, default decoder, enough
// chains to occupy the machine.)
static bool lane_schedule() { return env_int("LLICTI_DECODE_LANES", 1) != 0; }                // 0: the group schedule (A/B, tests)
static long long lane_min_warps() { return env_int("LLICTI_LANE_MIN_WARPS", 148); }
static bool group_scheduled(const llicti_config &cfg, const DecodeGeom &dg, int n) {
    if (cfg.sub_len <= 0 || cfg.decode_impl != 0) return false;
    if (lane_schedule()) return ((long long)n * dg.S + kLaneChains - 1) / kLaneChains >= lane_min_warps();
    return (long long)n * dg.S * group_lanes() / 32 >= group_min_warps();
}

// Window items the workspace must hold for batches of up to max_images images: the largest band that can reach a
// windowed schedule (torchac-compatible streams: every band; substream container: the bands with too few chains for
// the group schedule at some batch size n <= max_images).
int64_t decode_items_capacity(const llicti_config &cfg, const Plan &p, int max_images) {
    int64_t cap = 0;
    for (int s = 0; s < p.g.num_scales; ++s)
        for (int b = 0; b < 3; ++b) {
            const DecodeGeom dg = make_decode_geom(p, s, b);
            int n_win = max_images;                                  // largest batch that still takes windows for this band
            while (n_win > 0 && group_scheduled(cfg, dg, n_win)) --n_win;
            cap = std::max<int64_t>(cap, 3ll * n_win * dg.S * dg.items_per_chain);
        }
    return cap;
}

int read_decode_stats(uint64_t *out, int reset) {
    unsigned long long h[8];
    LLICTI_CUDA(cudaMemcpyFromSymbol(h, g_decode_stats, sizeof(h)));
    for (int i = 0; i < 8; ++i) out[i] = h[i];
    if (reset) {
        unsigned long long z[8] = {};
        LLICTI_CUDA(cudaMemcpyToSymbol(g_decode_stats, z, sizeof(z)));
    }
    return LLICTI_OK;
}

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

int launch_decode_band(llicti_ctx *ctx, const Plan &p, int scale, int band, const float *params, int16_t *planes,
                       const int32_t *minmax, int n, const uint8_t *blob, const uint64_t *suboff,
                       const uint32_t *sublen, cudaStream_t st) {
    const DecodeGeom dg = make_decode_geom(p, scale, band);
    const int total_sub = (int)p.g.substreams;
    if (ctx->cfg.decode_impl == 1) {
        ProfScope prof_(ctx, KC_DECODE, st);
        dim3 grid((dg.S + 3) / 4, n);
        decode_band_warp_kernel<<<grid, 128, 0, st>>>(params, planes, minmax, dg, ctx->num, blob, suboff, sublen, total_sub);
        ctx->launches += 1;
        LLICTI_CUDA(cudaGetLastError());
        return LLICTI_OK;
    }
    // Group schedule where the launch has enough chains to occupy the machine (its serial chain per symbol is long: the
    // whole CDF evaluation); with few chains -- the coarse scales -- the windows are produced by all SMs in parallel and
    // the serial part is the short chain kernel.
    const int G = group_lanes();
    if (group_scheduled(ctx->cfg, dg, n) && lane_schedule()) {
        ProfScope prof_(ctx, KC_DECODE, st);
        const long long warps = ((long long)n * dg.S + kLaneChains - 1) / kLaneChains;
        const int blocks = (int)((warps + 3) / 4);
        const int skew = env_int("LLICTI_TEST_GUESS_SKEW", 0);
        static const int occ = env_int("LLICTI_LANE_OCC", 6);      // A/B: 4 = the 118-register build
        // (the carve-out is a per-device attribute of the function: set at every launch, a host-side call of about a microsecond)
#define LLICTI_LANE_LAUNCH(OCC) do { \
            LLICTI_CUDA(cudaFuncSetAttribute(decode_band_lane_kernel<OCC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
            decode_band_lane_kernel<OCC><<<blocks, 128, 0, st>>>(params, planes, minmax, dg, ctx->num, blob, suboff, sublen, total_sub, n, skew); } while (0)
        if (occ >= 6) LLICTI_LANE_LAUNCH(6);
        else if (occ == 5) LLICTI_LANE_LAUNCH(5);
        else LLICTI_LANE_LAUNCH(4);
#undef LLICTI_LANE_LAUNCH
        ctx->launches += 1;
        LLICTI_CUDA(cudaGetLastError());
        return LLICTI_OK;
    }
    if (group_scheduled(ctx->cfg, dg, n)) {
        ProfScope prof_(ctx, KC_DECODE, st);
        const long long threads = (long long)n * dg.S * G;
        const int blocks = (int)((threads + 127) / 128);
        const bool locate = env_int("LLICTI_GROUP_LOCATE", 1) != 0;
#define LLICTI_GROUP_LAUNCH(GG, LL) decode_band_group_kernel<GG, LL><<<blocks, 128, 0, st>>>(params, planes, minmax, dg, ctx->num, blob, suboff, sublen, total_sub, n)
        if (G >= 16) { if (locate) LLICTI_GROUP_LAUNCH(16, true); else LLICTI_GROUP_LAUNCH(16, false); }
        else if (G >= 8) { if (locate) LLICTI_GROUP_LAUNCH(8, true); else LLICTI_GROUP_LAUNCH(8, false); }
        else if (G >= 4) { if (locate) LLICTI_GROUP_LAUNCH(4, true); else LLICTI_GROUP_LAUNCH(4, false); }
        else { if (locate) LLICTI_GROUP_LAUNCH(2, true); else LLICTI_GROUP_LAUNCH(2, false); }
#undef LLICTI_GROUP_LAUNCH
        ctx->launches += 1;
        LLICTI_CUDA(cudaGetLastError());
        return LLICTI_OK;
    }
    int sm_count = 0;
    {
        const int rc = device_sm_count(ctx, &sm_count);
        if (rc) return rc;
    }
    uint4 *items = reinterpret_cast<uint4 *>(ctx->d_items);
    int16_t *syms = ctx->d_syms;
    const size_t sym_cap = (size_t)ctx->sym_cap;
    const long long items_band = 3ll * n * dg.S * dg.items_per_chain;
    LLICTI_REQUIRE(items && syms && items_band <= ctx->items_cap && (size_t)dg.n_sym <= sym_cap, "decode workspace too small");
    // piped: a grid of sm_count x ctas_per_sm one-warp CTAs (32 resident per SM at most, so the grid
    // is co-resident); consumers claim ceil(3n / 4) SMs, the rest produce
    // consumers run fastest alone on an SM (measured: 1 per SM 54.8 ms, 2 per SM 59.4 ms on c1); give them up
    // to half of the SMs, the other half produces
    const int cons_auto = (3 * n + sm_count / 2 - 1) / (sm_count / 2);
    const int cons_per_sm = std::min(std::max(env_int("LLICTI_PIPE_CONS_PER_SM", cons_auto), 1), kConsPerSmMax);
    int &pipe_resident = ctx->pipe_resident;       // one-warp CTAs of the piped kernel that fit on one SM
    if (!pipe_resident)
        LLICTI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pipe_resident, decode_band_pipe_kernel, 32, 0));
    const int ctas_per_sm = std::min(std::min(std::max(env_int("LLICTI_PIPE_CTAS_PER_SM", 12), cons_per_sm + 1), 24), pipe_resident);
    const int cons_sms = (3 * n + cons_per_sm - 1) / cons_per_sm;
    const bool piped = !ctx->no_coresidency && dg.S == 1 && ctas_per_sm > cons_per_sm && cons_sms <= sm_count / 2 && 3 * n <= kConsPerSmMax * (sm_count / 2) && sm_count <= 256 &&
                       !env_int("LLICTI_NO_PIPE", 0);
    if (piped) {
        ProfScope prof_(ctx, KC_DECODE, st);
        // sentinels over the symbol arrays (data = flag), zeros over the role/ticket words and item flags
        LLICTI_CUDA(cudaMemsetAsync(syms, 0x80, (size_t)n * 3 * sym_cap * sizeof(int16_t), st));
        LLICTI_CUDA(cudaMemsetAsync(ctx->d_item_flags, 0, ((size_t)items_band + kCtlWords) * sizeof(uint32_t), st));
        decode_band_pipe_kernel<<<sm_count * ctas_per_sm, 32, 0, st>>>(params, syms, sym_cap, minmax, dg, ctx->num, items,
                                                                      ctx->d_item_flags, ctx->d_item_flags + kCtlWords, cons_per_sm, blob,
                                                                      suboff, sublen, total_sub, n);
        ctx->launches += 1;
    } else {
        // The window kernel is issue bound (erfc), the chain kernel latency bound: the batch is cut in two halves that
        // run one kernel out of phase on two streams, so the windows of one half overlap the chains of the other.  (Not
        // while the per-class profile is on: its events would time overlapping kernels.)
        const bool halves = n >= 2 && !ctx->prof_on && ctx->side_stream && ctx->ev_fork && ctx->ev_join && !env_int("LLICTI_NO_HALVES", 0);
        const int n_a = halves ? (n + 1) / 2 : n;
        const size_t P = (size_t)dg.Hs * dg.Ws;
        auto run_half = [&](int img0, int cnt, cudaStream_t s, bool fork_after_first) {
            const long long win_warps = (long long)cnt * dg.S * dg.items_per_chain;
            const int win_blocks = (int)std::min<long long>((win_warps + 3) / 4, (long long)sm_count * 16);
            const int con_blocks = (int)(((long long)cnt * dg.S + 3) / 4);
            const float *pr = params + (size_t)img0 * kParamCh * P;
            int16_t *sy = syms + (size_t)img0 * 3 * sym_cap;
            const int32_t *mm = minmax + (size_t)img0 * 4;
            uint4 *itm = items + (size_t)img0 * 3 * dg.S * dg.items_per_chain * kItemU4;
            const uint64_t *so = suboff + (size_t)img0 * total_sub;
            const uint32_t *sl = sublen + (size_t)img0 * total_sub;
            for (int clr = 0; clr < 3; ++clr) {
                {
                    ProfScope prof_(ctx, KC_WINDOW, s);
                    window_kernel<<<win_blocks, 128, 0, s>>>(pr, sy, sym_cap, mm, dg, clr, ctx->num, itm, cnt);
                }
                if (clr == 0 && fork_after_first) {
                    cudaEventRecord((cudaEvent_t)ctx->ev_fork, s);
                    cudaStreamWaitEvent((cudaStream_t)ctx->side_stream, (cudaEvent_t)ctx->ev_fork, 0);
                }
                {
                    ProfScope prof_(ctx, KC_DECODE, s);
                    consume_kernel<<<con_blocks, 128, 0, s>>>(pr, sy, sym_cap, mm, dg, clr, ctx->num, itm, blob, so, sl, total_sub, cnt);
                }
                ctx->launches += 2;
            }
        };
        run_half(0, n_a, st, halves);
        if (halves) {
            run_half(n_a, n - n_a, (cudaStream_t)ctx->side_stream, false);
            LLICTI_CUDA(cudaEventRecord((cudaEvent_t)ctx->ev_join, (cudaStream_t)ctx->side_stream));
            LLICTI_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)ctx->ev_join, 0));
        }
    }
    {
        ProfScope prof_(ctx, KC_MERGE, st);
        scatter_band_kernel<<<dim3((dg.n_sym + 255) / 256, 3 * n), 256, 0, st>>>(syms, sym_cap, planes, minmax, dg, n, 0, dg.n_sym);
        ctx->launches += 1;
    }
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

// ---- wavefront schedule, host side ------------------------------------------------------------
// The wavefront step needs two kernels that really run at the same time.  Tools that serialise
// kernel launches (Nsight Compute's kernel replay does) would make the consumer kernel wait for a
// producer kernel that cannot start, so the capability is probed once per context: a kernel waits a
// few milliseconds for a flag that only a second kernel on another stream can set.
__global__ void probe_wait_kernel(volatile uint32_t *flag, uint32_t *result) {
    const long long t0 = clock64();
    while (*flag == 0u && clock64() - t0 < 8000000ll) __nanosleep(1000);      // ~4 ms at 2 GHz
    *result = *flag;
}
__global__ void probe_set_kernel(volatile uint32_t *flag) { *flag = 1u; }

int probe_concurrent_kernels(llicti_ctx *ctx, bool *ok) {
    *ok = false;
    uint32_t *d = nullptr, h = 0;
    LLICTI_CUDA(cudaMalloc((void **)&d, 2 * sizeof(uint32_t)));
    cudaStream_t side = (cudaStream_t)ctx->side_stream, second = nullptr;
    LLICTI_CUDA(cudaStreamCreateWithFlags(&second, cudaStreamNonBlocking));
    // load every kernel that later runs next to a waiting kernel now (lazy module loading may have to
    // synchronise the context on a first launch), and run the probe pair once back to back
    cudaFuncAttributes fa;
    LLICTI_CUDA(cudaFuncGetAttributes(&fa, probe_wait_kernel));
    LLICTI_CUDA(cudaFuncGetAttributes(&fa, probe_set_kernel));
    LLICTI_CUDA(cudaFuncGetAttributes(&fa, wave_consume_kernel));
    LLICTI_CUDA(cudaFuncGetAttributes(&fa, wave_produce_kernel));
    LLICTI_CUDA(cudaFuncGetAttributes(&fa, scatter_band_kernel));
    LLICTI_CUDA(cudaMemset(d, 0, 2 * sizeof(uint32_t)));
    probe_set_kernel<<<1, 1, 0, second>>>(d);
    probe_wait_kernel<<<1, 1, 0, second>>>(d, d + 1);
    LLICTI_CUDA(cudaStreamSynchronize(second));
    LLICTI_CUDA(cudaMemset(d, 0, 2 * sizeof(uint32_t)));
    LLICTI_CUDA(cudaDeviceSynchronize());
    probe_wait_kernel<<<1, 1, 0, side>>>(d, d + 1);
    probe_set_kernel<<<1, 1, 0, second>>>(d);
    cudaError_t e = cudaStreamSynchronize(side);
    if (e == cudaSuccess) e = cudaStreamSynchronize(second);
    if (e == cudaSuccess) e = cudaMemcpy(&h, d + 1, sizeof(h), cudaMemcpyDeviceToHost);
    cudaStreamDestroy(second);
    cudaFree(d);
    if (e != cudaSuccess) { set_error("concurrency probe: %s", cudaGetErrorString(e)); return LLICTI_E_CUDA; }
    *ok = h != 0u;
    return LLICTI_OK;
}

static int env_int(const char *name, int dflt);
static int wave_strips(int Hs) {      // strips of >= 32 rows, at most 8 (measured optimum on 768x512: 4 -> 46.4 ms, 8 -> 40.5 ms, 16 -> 41.5 ms)
    return std::min(std::max(Hs / env_int("LLICTI_WAVE_STRIP_ROWS", 32), 1), env_int("LLICTI_WAVE_MAX_STRIPS", 8));
}

// Can the scale be decoded by the wavefront schedule?  (torchac-compatible streams, the tcgen05 CNN --
// the only one with row ranges --, enough rows for at least two strips, few enough chains for one
// consumer warp each, and a workspace reserved with three bands' worth of buffers.)
bool wave_eligible(const llicti_ctx *ctx, const Plan &p, int scale, int n) {
    if (ctx->cfg.decode_impl != 0 || ctx->cfg.cnn_impl != LLICTI_CNN_TCGEN05 || ctx->cfg.sub_len != 0) return false;
    if (!ctx->wave_ws || !ctx->concurrent_kernels || ctx->no_coresidency || env_int("LLICTI_NO_WAVE", 0) || env_int("LLICTI_NO_PIPE", 0)) return false;
    if (wave_strips(p.g.Hs[scale]) < 2) return false;
    // A strip is decoded in whole 32-symbol items, so band b may stop up to ceil(31 / crop_w) rows short of the strip's last
    // row; the three-row lag between the bands covers ONE missing row (launch_decode_scale_wave), i.e. rows of >= 32 symbols.
    for (int b = 0; b < 3; ++b)
        if (p.g.crop_w[scale][b] < 32) return false;
    return 9 * n <= kConsPerSmMax * 64;
}

int wave_bands_in_workspace(const llicti_config &cfg, int max_images) {
    // (the encoder's d_params holds one band; only the wavefront decode keeps three in flight)
    return (cfg.decode_impl == 0 && cfg.cnn_impl == LLICTI_CNN_TCGEN05 && cfg.sub_len == 0 && 9 * max_images <= kConsPerSmMax * 64) ? 3 : 1;
}

int launch_decode_scale_wave(llicti_ctx *ctx, const Plan &p, int scale, int16_t *planes, const int32_t *minmax, int n,
                             const uint8_t *blob, const uint64_t *suboff, const uint32_t *sublen, cudaStream_t st) {
    int sm_count = 0;
    {
        const int rc = device_sm_count(ctx, &sm_count);
        if (rc) return rc;
    }
    const int regs_per_sm = ctx->regs_per_sm;
    int &prod_resident = ctx->prod_resident, &cons_regs = ctx->cons_regs, &prod_regs = ctx->prod_regs;
    if (!prod_resident) {
        LLICTI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&prod_resident, wave_produce_kernel, 128, 0));
        cudaFuncAttributes fa;
        LLICTI_CUDA(cudaFuncGetAttributes(&fa, wave_consume_kernel));
        cons_regs = (fa.numRegs + 7) / 8 * 8;
        LLICTI_CUDA(cudaFuncGetAttributes(&fa, wave_produce_kernel));
        prod_regs = (fa.numRegs + 7) / 8 * 8;
    }
    LLICTI_REQUIRE(ctx->side_stream && ctx->ev_fork && ctx->ev_join, "wavefront decode: no side stream");
    cudaStream_t side = (cudaStream_t)ctx->side_stream;
    const int Hs = p.g.Hs[scale], Ws = p.g.Ws[scale];
    const int K = wave_strips(Hs);
    const size_t P0 = (size_t)p.g.Hs[0] * p.g.Ws[0];
    const size_t sym_cap = (size_t)ctx->sym_cap;
    const size_t nws = (size_t)ctx->ws_images;
    const int total_sub = (int)p.g.substreams;
    DecodeGeom dg[3];
    float *params[3];
    int16_t *syms[3];
    uint4 *items[3];
    uint32_t *flags[3];
    const size_t items_band = (size_t)ctx->items_cap;        // items per band in the workspace
    for (int b = 0; b < 3; ++b) {
        dg[b] = make_decode_geom(p, scale, b);
        LLICTI_REQUIRE(dg[b].S == 1 && 3ll * n * dg[b].items_per_chain <= ctx->items_cap && (size_t)dg[b].n_sym <= sym_cap,
                       "wavefront decode: workspace too small");
        params[b] = ctx->d_params + (size_t)b * nws * kParamCh * P0;
        syms[b] = ctx->d_syms + (size_t)b * nws * 3 * sym_cap;
        items[b] = reinterpret_cast<uint4 *>(ctx->d_items) + (size_t)b * items_band * kItemU4;
        flags[b] = ctx->d_item_flags + kCtlWords + (size_t)b * items_band;
    }
    // sentinels over the symbol arrays (data = flag), zeros over the item flags of the three bands
    LLICTI_CUDA(cudaMemsetAsync(ctx->d_syms, 0x80, 3 * nws * 3 * sym_cap * sizeof(int16_t), st));
    LLICTI_CUDA(cudaMemsetAsync(ctx->d_item_flags, 0, (kCtlWords + 3 * items_band) * sizeof(uint32_t), st));

    // producer CTAs (4 warps, 64 registers = 8 K registers each): `safe` per SM leave room for a consumer CTA (4 warps x
    // ~160 registers = 20 K registers) whichever kernel starts first; the CTAs beyond them stay only when the consumers
    // are already running (wave_produce_kernel)
    // chains per consumer CTA: one per scheduler is fastest (measured on c1, decode kernels per batch: 4 per CTA 31.0 ms,
    // 6: 33.2 ms, 8: 33.7 ms, 10: 36.7 ms -- a second chain on a scheduler slows both more than the freed SMs help)
    const int cpc = std::min(std::max(env_int("LLICTI_WAVE_CHAINS_PER_CTA", 4), 1), kWaveChainsMax);
    const int consumer_ctas = (9 * n + cpc - 1) / cpc;
    LLICTI_REQUIRE(consumer_ctas <= sm_count, "wavefront decode: more consumer CTAs than SMs");
    const int room = (regs_per_sm - cpc * 32 * cons_regs) / (128 * prod_regs);     // producer CTAs next to a consumer CTA (registers)
    const int room_thr = (2048 - cpc * 32) / 128;
    const int safe_per_sm = std::max(std::min(std::min(room, room_thr), prod_resident), 0);
    LLICTI_REQUIRE(safe_per_sm >= 1, "wavefront decode: no room for producers next to a consumer CTA");
    const int prod_per_sm = std::min(std::max(env_int("LLICTI_WAVE_PRODUCER_CTAS_PER_SM", prod_resident - 1), 1), std::max(prod_resident, 1));
    // Strip k of band b covers rows [row_of(b, k), row_of(b, k+1)): the strips of band b+1 end three rows before
    // those of band b.  The CNN of band b+1 reads band b up to two rows below the row it evaluates, and a strip is
    // decoded in whole 32-symbol items (up to one row short of its last row), so strip k of band b+1 depends on
    // strips <= k of band b only: the bands run ONE time step apart.
    auto row_of = [&](int b, int k) {
        if (k <= 0) return 0;
        if (k >= K) return Hs;
        return std::max((int)((long long)Hs * k / K) - 3 * b, 0);
    };
    auto item_of = [&](int b, int k) {       // items of band b complete after strip k-1
        if (k >= K) return dg[b].items_per_chain;
        return (int)std::min<long long>((long long)std::min(row_of(b, k), dg[b].crop_h) * dg[b].crop_w / 32, dg[b].items_per_chain);
    };
    for (int T = 0; T < K + 2; ++T) {
        WaveArgs wa;
        bool any = false;
        for (int b = 0; b < 3; ++b) {
            const int s = T - b;
            wa.b[b].dg = dg[b];
            wa.b[b].params = params[b]; wa.b[b].syms = syms[b]; wa.b[b].items = items[b]; wa.b[b].flags = flags[b];
            wa.b[b].it0 = wa.b[b].it1 = 0;
            if (s < 0 || s >= K) continue;
            // network outputs of the rows this strip decodes (band 0 depends on x00 only: all rows at once)
            int rc = LLICTI_OK;
            if (b == 0) { if (s == 0) rc = launch_cnn_tc(ctx, 0, planes, n, Hs, Ws, params[0], st); }
            else if (row_of(b, s + 1) > row_of(b, s))
                rc = launch_cnn_tc(ctx, b, planes, n, Hs, Ws, params[b], st, row_of(b, s), row_of(b, s + 1) - row_of(b, s));
            if (rc) return rc;
            wa.b[b].it0 = item_of(b, s);
            wa.b[b].it1 = item_of(b, s + 1);
            any |= wa.b[b].it1 > wa.b[b].it0;
        }
        if (any) {
            ProfScope prof_(ctx, KC_DECODE, st);
            // (Measured and dropped: producing the Y windows of the strips ahead of the pair with the whole GPU.  With all
            // three channels on 94 SMs the producers do not keep up with 216 chains -- the Y chains alone wait 0.47 of
            // 1.8 ms per step -- and produced ahead the waiting goes (average chain 1.78 -> 1.36 ms); but a step lasts as
            // long as its slowest chain, the image with the most redone chunks, and the extra launches cost more than
            // they saved: 29.0 -> 31.2 ms per c1 batch.)
            LLICTI_CUDA(cudaMemsetAsync(ctx->d_item_flags, 0, kCtlWords * sizeof(uint32_t), st));
            // fork: consumers on the caller's stream (they start first), producers on the side stream; join before the scatter
            LLICTI_CUDA(cudaEventRecord((cudaEvent_t)ctx->ev_fork, st));
            LLICTI_CUDA(cudaStreamWaitEvent(side, (cudaEvent_t)ctx->ev_fork, 0));
            wave_consume_kernel<<<consumer_ctas, 32 * cpc, 0, st>>>(wa, minmax, ctx->num, sym_cap, ctx->d_item_flags, blob, suboff,
                                                              sublen, total_sub, n,
                                                              reinterpret_cast<ChainState *>(ctx->d_chain_state_raw));
            wave_produce_kernel<<<sm_count * prod_per_sm, 128, 0, side>>>(wa, minmax, ctx->num, sym_cap, ctx->d_item_flags, n,
                                                                         sm_count * std::min(safe_per_sm, prod_per_sm), consumer_ctas, env_int("LLICTI_WAVE_SHARE_SMS", 0),
                                                                         std::max(env_int("LLICTI_WAVE_PATTERN", 122), 1));
            LLICTI_CUDA(cudaEventRecord((cudaEvent_t)ctx->ev_join, side));
            LLICTI_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)ctx->ev_join, 0));
            ctx->launches += 2;
            if (env_int("LLICTI_WAVE_DEBUG", 0)) wave_dbg_kernel<<<1, 1, 0, st>>>(T, 9 * n);
        }
        for (int b = 0; b < 3; ++b) {
            if (wa.b[b].it1 <= wa.b[b].it0) continue;
            ProfScope prof_(ctx, KC_MERGE, st);
            const int sym0 = wa.b[b].it0 * 32, sym1 = std::min(wa.b[b].it1 * 32, dg[b].n_sym);
            scatter_band_kernel<<<dim3((sym1 - sym0 + 255) / 256, 3 * n), 256, 0, st>>>(syms[b], sym_cap, planes, minmax, dg[b], n,
                                                                                      sym0, sym1);
            ctx->launches += 1;
        }
    }
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
