// GMM-CDF bounds, arithmetic coding / decoding over interleaved substreams, and the stream
// container (compaction on encode, indexing on decode).
//
// Reference path replaced: LLICTIEntropyLayer.compress / decompress inner loops
// (graphs/models/LLICTI_nets.py:378-411 and :463-498): per (scale, band, colour channel)
// mean coupling -> get_cdfs -> torchac.  The reference materialises H x W x Lp int16 tables
// and ships them to a single CPU thread; here nothing but 4 bytes per symbol (encode) or
// the 60 network outputs per position (decode) ever leave the chip's caches.
#include "common.cuh"

#include <cstdlib>
#include "gmm.cuh"
#include "rangecoder.cuh"

namespace llicti {

constexpr long long kWarpChainLimit = 148 * 16;   // chains up to which the coder runs one warp per chain

// ------------------------------------------------------------------------------------------
// Dense table / flat bounds for one stream (parity entry points llicti_cdf_table / _bounds)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
cdf_table_kernel(const float *__restrict__ params, const int16_t *__restrict__ yband, int clr, int min_val,
                 int max_val, int P, NumericsProfile np, int16_t *__restrict__ table) {
    // one warp per position: lanes stride over the Lp table entries
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= P) return;
    GmmChannel c;
    load_channel(params, (size_t)P, (size_t)warp, clr, yband[warp], yband[(size_t)P + warp], np, c);
    const CdfGrid g = make_grid(min_val, max_val);
    int16_t *row = table + (size_t)warp * g.Lp;
    for (int k = lane; k < g.Lp; k += 32) row[k] = (int16_t)cdf_q(c, g, k, np);
}

__device__ __forceinline__ uint32_t symbol_bounds(const GmmChannel &c, const CdfGrid &g, int sym,
                                                  const NumericsProfile &np) {
    const uint32_t c_low = cdf_q(c, g, sym, np);
    const uint32_t c_high = (sym == g.Lp - 2) ? 0x10000u : cdf_q(c, g, sym + 1, np);
    return c_low | ((c_high - 1u) << 16);
}

__global__ void __launch_bounds__(128)
cdf_bounds_flat_kernel(const float *__restrict__ params, const int16_t *__restrict__ yband, int clr, int min_val,
                       int max_val, int P, NumericsProfile np, uint32_t *__restrict__ bounds) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    GmmChannel c;
    load_channel(params, (size_t)P, (size_t)i, clr, yband[i], yband[(size_t)P + i], np, c);
    const CdfGrid g = make_grid(min_val, max_val);
    const int sym = (int)yband[(size_t)clr * P + i] - min_val;
    bounds[i] = symbol_bounds(c, g, sym, np);
}

// ------------------------------------------------------------------------------------------
// Encode side: bounds of a whole band (all images, 3 colour channels per thread)
// ------------------------------------------------------------------------------------------
struct BandGeom {
    int Hs, Ws, crop_h, crop_w, band;
    int64_t sym_off[3];    // first symbol of the Y / Co / Cg stream among the image's symbols
};

__global__ void __launch_bounds__(128, 8)
band_bounds_kernel(const float *__restrict__ params, const int16_t *__restrict__ planes,
                   const int32_t *__restrict__ minmax, BandGeom bg, NumericsProfile np,
                   uint32_t *__restrict__ bounds, int64_t sym_stride) {
    // the three sampling grids of the image (two fp64 divisions each): once per block, not per position
    __shared__ CdfGrid s_grid[3];
    const int img = blockIdx.y;
    if (threadIdx.x < 3) {
        const int32_t *mm = minmax + img * 4;
        s_grid[threadIdx.x] = threadIdx.x == 0 ? make_grid(-127, 128) : threadIdx.x == 1 ? make_grid(mm[0], mm[2]) : make_grid(mm[1], mm[3]);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // index in the cropped raster
    const int n_sym = bg.crop_h * bg.crop_w;
    if (i >= n_sym) return;
    const int r = i / bg.crop_w, c = i - r * bg.crop_w;
    const size_t P = (size_t)bg.Hs * bg.Ws;
    const size_t pidx = (size_t)r * bg.Ws + c;
    const float *pp = params + (size_t)img * kParamCh * P;
    const int16_t *yb = planes + (size_t)img * 12 * P + (size_t)(3 * (bg.band + 1)) * P + pidx;
    const int y0 = yb[0], y1 = yb[P], y2 = yb[2 * P];
    uint32_t *out = bounds + (size_t)img * sym_stride + i;
    // one copy of the channel's code, run three times (the unrolled form is 3500 instructions: the launch stalled on
    // instruction fetch, ncu "no instruction" 3.9 warps per issue)
#pragma unroll 1
    for (int clr = 0; clr < 3; ++clr) {
        GmmChannel ch;
        load_channel(pp, P, pidx, clr, y0, y1, np, ch);
        const CdfGrid g = s_grid[clr];
        const int yv = clr == 0 ? y0 : clr == 1 ? y1 : y2;
        const int64_t so = clr == 0 ? bg.sym_off[0] : clr == 1 ? bg.sym_off[1] : bg.sym_off[2];
        out[so] = symbol_bounds(ch, g, yv - g.min_val, np);
    }
}

// ------------------------------------------------------------------------------------------
// Encode side: one thread per (image, stream, substream)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_stream(const StreamDesc *sd, int n_streams, int item) {
    int k = 0;
    while (k + 1 < n_streams && sd[k + 1].sub_first <= item) ++k;
    return k;
}

// Code symbols j, j+S, j+2S, ... ; eight bounds are fetched ahead of the serial coder chain so
// that the loads' latency overlaps the arithmetic of the previous eight symbols.
__device__ __forceinline__ void encode_strided(AcEncoder &enc, const uint32_t *__restrict__ b, int j, int n_sym, int S) {
    constexpr int U = 8;
    for (int i = j; i < n_sym; i += U * S) {
        uint32_t v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = i + u * S;
            v[u] = idx < n_sym ? b[idx] : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * S < n_sym) enc.encode(v[u] & 0xFFFFu, (v[u] >> 16) + 1u);
    }
}

// ------------------------------------------------------------------------------------------
// Warp-per-chain encoder for few, long chains (torchac-compatible streams).
//
// The only serial part of arithmetic ENcoding is the interval recurrence (low, high) -> (low,
// high); which bits a step emits depends on that step alone plus the count of pending underflow
// bits.  So a block of 32 symbols runs in two phases:
//   A  serial, 32 steps: the recurrence only (~20 instructions per symbol); lane s keeps step
//      s's record (the n bits it shifts out, n, the underflow count k);
//   B  parallel, one step per lane: pending counts by a segmented prefix sum, bit offsets by a
//      prefix sum, then every lane ORs its bit string into a shared staging buffer, and the
//      completed 32-bit words go out with one coalesced store.
// The bit stream is identical to torchac's bit-plus-follow output.
// ------------------------------------------------------------------------------------------
constexpr int kStageWords = 96;          // staging capacity per warp (3072 bits); see block_fits()

// clz of a non-zero word through the exponent of a round-toward-zero int -> float conversion (I2F: ~10 cycles;
// FLO: ~22 cycles on sm_100)
__device__ __forceinline__ int enc_clz_nz(uint32_t x) { return 158 - (int)(__float_as_uint(__uint2float_rz(x)) >> 23); }

struct WarpEncoder {
    uint32_t low, high, pending;
    uint32_t *stage;       // shared, kStageWords + 2 words, zero except for emitted bits
    uint32_t *out;         // 4-byte aligned slot
    uint32_t cap_words, nwords, carry;   // slot capacity, words written, valid bits in stage[0..] not yet written
    int overflow, lane;

    __device__ __forceinline__ void init(uint8_t *slot, uint32_t cap_bytes, uint32_t *stage_, int lane_) {
        low = 0; high = 0xFFFFFFFFu; pending = 0;
        stage = stage_; out = reinterpret_cast<uint32_t *>(slot); cap_words = cap_bytes / 4; nwords = 0; carry = 0;
        overflow = 0; lane = lane_;
        for (int w = lane; w < kStageWords + 2; w += 32) stage[w] = 0u;
        __syncwarp();
    }
    // OR `cnt` (<= 32) bits, right aligned in val, at bit position pos (MSB first)
    __device__ __forceinline__ void or_bits(uint32_t pos, uint32_t val, int cnt) {
        if (cnt <= 0) return;
        const uint64_t t = (((uint64_t)val) << (64 - cnt)) >> (pos & 31u);
        const uint32_t hi = (uint32_t)(t >> 32), lo = (uint32_t)t;
        if (hi) atomicOr(&stage[pos >> 5], hi);
        if (lo) atomicOr(&stage[(pos >> 5) + 1], lo);
    }
    __device__ __forceinline__ void or_ones(uint32_t pos, uint32_t cnt) {
        while (cnt > 0) {
            const uint32_t c = min(cnt, 32u);
            or_bits(pos, 0xFFFFFFFFu >> (32 - c), (int)c);
            pos += c;
            cnt -= c;
        }
    }
    // Write out the completed words of stage[] (bits [0, carry + added)), keep the partial one.
    __device__ __forceinline__ void flush(uint32_t added) {
        __syncwarp();
        const uint32_t bits = carry + added, nw = bits >> 5;
        const uint32_t partial = stage[nw];
        for (uint32_t w = lane; w < nw; w += 32) {
            if (nwords + w < cap_words) out[nwords + w] = __byte_perm(stage[w], 0, 0x0123);
        }
        if (nwords + nw > cap_words) overflow = 1;
        __syncwarp();
        for (uint32_t w = lane; w <= nw + 1; w += 32) stage[w] = w == 0 ? partial : 0u;
        nwords += nw;
        carry = bits & 31u;
        __syncwarp();
    }
    // One emitting step, written by lane 0 alone with flushes as needed (rare paths and finish()).
    __device__ __forceinline__ void emit_serial(uint32_t b, uint32_t run, uint32_t rest, int rest_cnt) {
        uint32_t added = 0;
        if (lane == 0) or_bits(carry, b, 1);
        added = 1;
        while (run > 0) {
            const uint32_t c = min(run, 1024u);
            if (lane == 0 && !b) or_ones(carry + added, c);     // the run repeats the complement of b
            added += c;
            run -= c;
            flush(added);
            added = 0;
        }
        if (lane == 0) or_bits(carry + added, rest, rest_cnt);
        added += rest_cnt;
        flush(added);
    }

    // 32 (or fewer) symbols: bounds in `cur` (one per lane), m valid.
    __device__ __forceinline__ void encode_block(uint32_t cur, int m) {
        uint32_t rec_nl, rec_nh;
        recur_block(cur, m, rec_nl, rec_nh);
        emit_block(rec_nl, rec_nh, m);
    }

    // ---- phase A: the interval recurrence, nothing else (lane s keeps (nl, nh) of step s); uses low / high only ----
    __device__ __forceinline__ void recur_block(uint32_t cur, int m, uint32_t &rec_nl, uint32_t &rec_nh) {
        // 32-bit arithmetic: (span * c) >> 16 is the high word of span * (c << 16) for span < 2^32; the full range
        // (span = 2^32, seen as 0) gives c << 16 itself, and c_high = 2^16 (c << 16 seen as 0) gives span.
        const uint32_t my_cl16 = cur << 16, my_ch16 = ((cur >> 16) + 1u) << 16;   // bounds unpacked in parallel, off the chain
        rec_nl = 0; rec_nh = 0xFFFFFFFFu;
        auto step = [&](int s, uint32_t cl16, uint32_t ch16) {
            const uint32_t span = high - low + 1u;
            const bool full = span == 0u;
            const uint32_t tl = full ? cl16 : __umulhi(span, cl16);
            const uint32_t th = ch16 == 0u ? span : full ? ch16 : __umulhi(span, ch16);
            const uint32_t nl = low + tl, nh = low + th - 1u;
            const int sh = enc_clz_nz((nl ^ nh) & ~((nl & ~nh) << 1));     // equal leading bits + underflow run (AcEncoder::encode)
            if (lane == s) { rec_nl = nl; rec_nh = nh; }
            low = (nl << sh) & 0x7FFFFFFFu;
            high = (nh << sh) | ~(0xFFFFFFFFu << sh) | 0x80000000u;
        };
        if (m == 32) {
            // the bounds of eight steps are broadcast while the previous eight run: no shuffle latency on the chain
            uint32_t cl[8], ch[8], ncl[8], nch[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { cl[k] = __shfl_sync(0xffffffffu, my_cl16, k); ch[k] = __shfl_sync(0xffffffffu, my_ch16, k); }
#pragma unroll
            for (int b8 = 0; b8 < 4; ++b8) {
                if (b8 < 3) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        ncl[k] = __shfl_sync(0xffffffffu, my_cl16, 8 * b8 + 8 + k);
                        nch[k] = __shfl_sync(0xffffffffu, my_ch16, 8 * b8 + 8 + k);
                    }
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) step(8 * b8 + k, cl[k], ch[k]);
#pragma unroll
                for (int k = 0; k < 8; ++k) { cl[k] = ncl[k]; ch[k] = nch[k]; }
            }
        } else {
#pragma unroll 1
            for (int s = 0; s < m; ++s) step(s, __shfl_sync(0xffffffffu, my_cl16, s), __shfl_sync(0xffffffffu, my_ch16, s));
        }
    }

    // ---- phase B: bit output of the 32 steps in parallel; uses pending and the staging buffer only ----
    __device__ __forceinline__ void emit_block(uint32_t rec_nl, uint32_t rec_nh, int m) {
        // per-step record, derived in parallel: the n bits shifted out, n, the underflow count k
        int rec_n = 0, rec_k = 0;
        uint32_t rec_bits = 0;
        if (lane < m) {
            const uint32_t d = rec_nl ^ rec_nh;
            rec_n = __clz(d);
            rec_k = __clz(d & ~((rec_nl & ~rec_nh) << 1)) - rec_n;
            rec_bits = __funnelshift_l(rec_nl, 0u, rec_n);
        }
        int kx = rec_k;                                                   // inclusive scan of k
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, kx, o); if (lane >= o) kx += y; }
        const int ktot = __shfl_sync(0xffffffffu, kx, 31);
        kx -= rec_k;                                                      // exclusive
        const uint32_t emask = __ballot_sync(0xffffffffu, rec_n > 0);
        const uint32_t prev = emask & ((1u << lane) - 1u);
        const int kx_prev = __shfl_sync(0xffffffffu, kx, prev ? 31 - __clz(prev) : 0);
        const uint32_t P = prev ? (uint32_t)(kx - kx_prev) : pending + (uint32_t)kx;   // pending bits when this step emits
        const uint32_t L = rec_n > 0 ? (uint32_t)rec_n + P : 0u;
        uint32_t ox = L;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, ox, o); if (lane >= o) ox += y; }
        const uint32_t ltot = __shfl_sync(0xffffffffu, ox, 31);
        ox -= L;
        const int kx_last = __shfl_sync(0xffffffffu, kx, emask ? 31 - __clz(emask) : 0);
        const uint32_t pending_out = emask ? (uint32_t)(ktot - kx_last) : pending + (uint32_t)ktot;
        if (carry + ltot <= (uint32_t)kStageWords * 32u) {
            if (rec_n > 0) {
                const uint32_t pos = carry + ox;
                if (P == 0) {
                    or_bits(pos, rec_bits, rec_n);
                } else {
                    const uint32_t b = rec_bits >> (rec_n - 1);
                    or_bits(pos, b, 1);
                    if (!b) or_ones(pos + 1, P);
                    or_bits(pos + 1 + P, rec_bits & ~(1u << (rec_n - 1)), rec_n - 1);
                }
            }
            flush(ltot);
        } else {
            // more pending bits than the staging buffer holds (astronomically rare): step by step
            for (int s = 0; s < 32; ++s) {
                const int n_s = __shfl_sync(0xffffffffu, rec_n, s);
                const uint32_t bits_s = __shfl_sync(0xffffffffu, rec_bits, s), P_s = __shfl_sync(0xffffffffu, P, s);
                if (n_s > 0) emit_serial(bits_s >> (n_s - 1), P_s, bits_s & ~(1u << (n_s - 1)), n_s - 1);
            }
        }
        pending = pending_out;
    }

    __device__ __forceinline__ uint32_t finish() {
        pending += 1;
        const uint32_t b = low < 0x40000000u ? 0u : 1u;
        emit_serial(b, pending, 0u, 0);
        // zero-pad to a byte boundary: carry (< 32) bits are left in stage[0]
        const uint32_t nb = (carry + 7) >> 3;
        const uint32_t word = stage[0];
        uint8_t *tail = reinterpret_cast<uint8_t *>(out + nwords);
        if (lane == 0)
            for (uint32_t i = 0; i < nb; ++i) {
                if (nwords * 4 + i < cap_words * 4) tail[i] = (uint8_t)(word >> (24 - 8 * i));
                else overflow = 1;
            }
        overflow = __shfl_sync(0xffffffffu, overflow, 0) | overflow;
        return nwords * 4 + nb;
    }
};

__device__ __forceinline__ void encode_strided_warp(WarpEncoder &enc, const uint32_t *__restrict__ b, int j, int n_sym, int S) {
    const int lane = enc.lane;
    const int n_steps = (n_sym - j + S - 1) / S;
    uint32_t cur = lane < n_steps ? b[(size_t)j + (size_t)lane * S] : 0u;
    for (int t0 = 0; t0 < n_steps; t0 += 32) {
        const int t1 = t0 + 32 + lane;
        const uint32_t nxt = t1 < n_steps ? b[(size_t)j + (size_t)t1 * S] : 0u;   // a block ahead of the coder
        enc.encode_block(cur, min(32, n_steps - t0));
        cur = nxt;
    }
}

__global__ void __launch_bounds__(128)
encode_all_warp_kernel(const StreamDesc *__restrict__ sd_g, int n_streams, int total_sub,
                       const uint32_t *__restrict__ bounds, int64_t sym_stride, uint8_t *__restrict__ scratch,
                       int64_t scratch_stride, uint32_t *__restrict__ sublen, int32_t *__restrict__ status) {
    __shared__ StreamDesc sd[kMaxStreams];
    __shared__ uint32_t stage[4][kStageWords + 2];
    for (int e = threadIdx.x; e < n_streams; e += blockDim.x) sd[e] = sd_g[e];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int img = blockIdx.y;
    if (item >= total_sub) return;
    const int k = find_stream(sd, n_streams, item);
    const StreamDesc &d = sd[k];
    const int j = item - d.sub_first;
    const uint32_t *b = bounds + (size_t)img * sym_stride + d.sym_off;
    WarpEncoder enc;
    enc.init(scratch + (size_t)img * scratch_stride + d.slot_off + (size_t)j * d.slot_bytes, (uint32_t)d.slot_bytes,
             stage[threadIdx.x >> 5], lane);
    encode_strided_warp(enc, b, j, d.n_sym, d.S);
    const uint32_t nb = enc.finish();
    if (lane == 0) {
        sublen[(size_t)img * total_sub + item] = nb;
        if (enc.overflow) atomicExch(status, LLICTI_E_NOMEM);
    }
}

// Two warps per chain: warp A runs the recurrence of block k + 1 while warp B packs the bits of block k.  The
// per-step records (nl, nh) travel through a double buffer in shared memory, handed over with named barriers
// (bar.arrive by the writer, bar.sync by the reader; 64 = both warps).  Two chains per 128-thread CTA:
// warps 0, 1 recur, warps 2, 3 pack; barriers 1 + 4 pair + {0, 1} = full[buf], + {2, 3} = empty[buf].
__device__ __forceinline__ void pair_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void pair_arrive(int id) { asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(128)
encode_all_pair_kernel(const StreamDesc *__restrict__ sd_g, int n_streams, int total_sub,
                       const uint32_t *__restrict__ bounds, int64_t sym_stride, uint8_t *__restrict__ scratch,
                       int64_t scratch_stride, uint32_t *__restrict__ sublen, int32_t *__restrict__ status) {
    __shared__ StreamDesc sd[kMaxStreams];
    __shared__ uint32_t stage[2][kStageWords + 2];
    __shared__ uint2 rec[2][2][32];           // [pair][buffer][step]
    __shared__ uint32_t final_low[2];
    for (int e = threadIdx.x; e < n_streams; e += blockDim.x) sd[e] = sd_g[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp & 1, packer = warp >> 1;
    const int item = blockIdx.x * 2 + pair;
    const int img = blockIdx.y;
    if (item >= total_sub) return;            // both warps of the pair leave
    const int k = find_stream(sd, n_streams, item);
    const StreamDesc &d = sd[k];
    const int j = item - d.sub_first;
    const int n_steps = (d.n_sym - j + d.S - 1) / d.S;
    const int n_blocks = (n_steps + 31) >> 5;
    const int bar0 = 1 + 4 * pair;
    WarpEncoder enc;
    enc.init(scratch + (size_t)img * scratch_stride + d.slot_off + (size_t)j * d.slot_bytes, (uint32_t)d.slot_bytes,
             stage[pair], lane);
    if (!packer) {
        const uint32_t *b = bounds + (size_t)img * sym_stride + d.sym_off;
        uint32_t cur = lane < n_steps ? b[(size_t)j + (size_t)lane * d.S] : 0u;
        for (int kb = 0; kb < n_blocks; ++kb) {
            const int t1 = kb * 32 + 32 + lane;
            const uint32_t nxt = t1 < n_steps ? b[(size_t)j + (size_t)t1 * d.S] : 0u;   // a block ahead of the coder
            uint32_t nl, nh;
            enc.recur_block(cur, min(32, n_steps - kb * 32), nl, nh);
            if (kb >= 2) pair_sync(bar0 + 2 + (kb & 1));          // the packer has read block kb - 2 out of this buffer
            rec[pair][kb & 1][lane] = make_uint2(nl, nh);
            if (kb == n_blocks - 1 && lane == 0) final_low[pair] = enc.low;
            pair_arrive(bar0 + (kb & 1));
            cur = nxt;
        }
        return;
    }
    for (int kb = 0; kb < n_blocks; ++kb) {
        pair_sync(bar0 + (kb & 1));
        const uint2 r = rec[pair][kb & 1][lane];
        if (kb == n_blocks - 1) enc.low = final_low[pair];
        if (kb + 2 < n_blocks) pair_arrive(bar0 + 2 + (kb & 1));
        enc.emit_block(r.x, r.y, min(32, n_steps - kb * 32));
    }
    const uint32_t nb = enc.finish();
    if (lane == 0) {
        sublen[(size_t)img * total_sub + item] = nb;
        if (enc.overflow) atomicExch(status, LLICTI_E_NOMEM);
    }
}

__global__ void __launch_bounds__(128)
encode_flat_warp_kernel(const uint32_t *__restrict__ bounds, int n_sym, int S, uint8_t *__restrict__ out, int slot_bytes,
                        uint32_t *__restrict__ lens, int32_t *__restrict__ status) {
    __shared__ uint32_t stage[4][kStageWords + 2];
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= S) return;
    WarpEncoder enc;
    enc.init(out + (size_t)j * slot_bytes, (uint32_t)slot_bytes, stage[threadIdx.x >> 5], lane);
    encode_strided_warp(enc, bounds, j, n_sym, S);
    const uint32_t nb = enc.finish();
    if (lane == 0) {
        lens[j] = nb;
        if (enc.overflow) atomicExch(status, LLICTI_E_NOMEM);
    }
}

__global__ void __launch_bounds__(128)
encode_all_kernel(const StreamDesc *__restrict__ sd_g, int n_streams, int total_sub,
                  const uint32_t *__restrict__ bounds, int64_t sym_stride, uint8_t *__restrict__ scratch,
                  int64_t scratch_stride, uint32_t *__restrict__ sublen, int32_t *__restrict__ status) {
    __shared__ StreamDesc sd[kMaxStreams];
    for (int e = threadIdx.x; e < n_streams; e += blockDim.x) sd[e] = sd_g[e];
    __syncthreads();
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    const int img = blockIdx.y;
    if (item >= total_sub) return;
    const int k = find_stream(sd, n_streams, item);
    const StreamDesc &d = sd[k];
    const int j = item - d.sub_first;
    const uint32_t *b = bounds + (size_t)img * sym_stride + d.sym_off;
    AcEncoder enc;
    enc.init(scratch + (size_t)img * scratch_stride + d.slot_off + (size_t)j * d.slot_bytes, (uint32_t)d.slot_bytes);
    encode_strided(enc, b, j, d.n_sym, d.S);
    const uint32_t nb = enc.finish();
    sublen[(size_t)img * total_sub + item] = nb;
    if (enc.bw.overflow) atomicExch(status, LLICTI_E_NOMEM);
}

__global__ void __launch_bounds__(128)
encode_flat_kernel(const uint32_t *__restrict__ bounds, int n_sym, int S, uint8_t *__restrict__ out, int slot_bytes,
                   uint32_t *__restrict__ lens, int32_t *__restrict__ status) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    AcEncoder enc;
    enc.init(out + (size_t)j * slot_bytes, (uint32_t)slot_bytes);
    encode_strided(enc, bounds, j, n_sym, S);
    lens[j] = enc.finish();
    if (enc.bw.overflow) atomicExch(status, LLICTI_E_NOMEM);
}

// ------------------------------------------------------------------------------------------
// Container: [u16 S][u16 len x S][payloads] per stream when sub_len > 0, raw torchac bytes else
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
stream_sizes_kernel(const StreamDesc *__restrict__ sd, int n_streams, int total_sub, int sub_mode,
                    const uint32_t *__restrict__ sublen, uint64_t *__restrict__ stream_bytes) {
    const int k = blockIdx.x, img = blockIdx.y;
    const StreamDesc d = sd[k];
    const uint32_t *l = sublen + (size_t)img * total_sub + d.sub_first;
    uint32_t s = 0;
    for (int j = threadIdx.x; j < d.S; j += blockDim.x) s += l[j];
    __shared__ uint32_t red[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0)
        stream_bytes[(size_t)img * n_streams + k] = (uint64_t)(red[0] + red[1] + red[2] + red[3]) + (sub_mode ? 2 + 2 * d.S : 0);
}

// Exclusive scan of `count` uint64 values by one CTA (count = images * streams, a few 10^4 at most).
__global__ void __launch_bounds__(1024)
scan_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int count) {
    __shared__ uint64_t warp_tot[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < count; base += 1024) {
        const int i = base + threadIdx.x;
        const uint64_t v = i < count ? in[i] : 0;
        uint64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint64_t t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            warp_tot[lane] = t;
        }
        __syncthreads();
        const uint64_t prefix = carry + (wid ? warp_tot[wid - 1] : 0) + x - v;
        if (i < count) out[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[count] = carry;
}

// One CTA per (stream, image): write the container header and copy the substream payloads.
__global__ void __launch_bounds__(256)
gather_kernel(const StreamDesc *__restrict__ sd, int n_streams, int total_sub, int sub_mode,
              const uint8_t *__restrict__ scratch, int64_t scratch_stride, const uint32_t *__restrict__ sublen,
              const uint64_t *__restrict__ stream_off, uint8_t *__restrict__ out, uint64_t out_cap,
              int32_t *__restrict__ status) {
    const int k = blockIdx.x, img = blockIdx.y;
    const StreamDesc d = sd[k];
    const uint32_t *l = sublen + (size_t)img * total_sub + d.sub_first;
    const uint64_t o0 = stream_off[(size_t)img * n_streams + k];
    const uint64_t o1 = stream_off[(size_t)img * n_streams + k + 1];
    if (o1 > out_cap) {
        if (threadIdx.x == 0) atomicExch(status, LLICTI_E_NOMEM);
        return;
    }
    uint8_t *dst = out + o0;
    const uint8_t *src = scratch + (size_t)img * scratch_stride + d.slot_off;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t hdr = 0;
    if (sub_mode) {
        hdr = 2 + 2 * d.S;
        if (threadIdx.x == 0) { dst[0] = (uint8_t)(d.S & 0xFF); dst[1] = (uint8_t)(d.S >> 8); }
        for (int j = threadIdx.x; j < d.S; j += blockDim.x) {
            const uint32_t v = l[j];
            if (v > 0xFFFFu) atomicExch(status, LLICTI_E_NOMEM);
            dst[2 + 2 * j] = (uint8_t)(v & 0xFF);
            dst[3 + 2 * j] = (uint8_t)(v >> 8);
        }
    }
    // each warp walks the substreams in order, keeping a running payload offset
    __shared__ uint32_t chunk_off[256];
    uint32_t running = hdr;
    for (int base = 0; base < d.S; base += 256) {
        const int j = base + threadIdx.x;
        const uint32_t v = j < d.S ? l[j] : 0;
        // block exclusive scan of 256 values
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        __shared__ uint32_t wt[8];
        if (lane == 31) wt[wid] = x;
        __syncthreads();
        uint32_t wp = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < wid) wp += wt[w];
            tot += wt[w];
        }
        chunk_off[threadIdx.x] = running + wp + x - v;
        __syncthreads();
        const int cnt = min(256, d.S - base);
        for (int jj = wid; jj < cnt; jj += 8) {
            const uint32_t len = l[base + jj];
            const uint8_t *s = src + (size_t)(base + jj) * d.slot_bytes;
            uint8_t *t = dst + chunk_off[jj];
            for (uint32_t b = lane; b < len; b += 32) t[b] = s[b];
        }
        running += tot;
        __syncthreads();
    }
}

// Decode side: parse the per-stream containers into absolute substream offsets / lengths.
__global__ void __launch_bounds__(256)
index_streams_kernel(const StreamDesc *__restrict__ sd, int n_streams, int total_sub, int sub_mode,
                     const uint8_t *__restrict__ blob, uint64_t blob_bytes, int n_images, const uint64_t *__restrict__ stream_off,
                     uint64_t *__restrict__ suboff, uint32_t *__restrict__ sublen, int32_t *__restrict__ status) {
    const int k = blockIdx.x, img = blockIdx.y;
    const StreamDesc d = sd[k];
    const uint64_t o0 = stream_off[(size_t)img * n_streams + k];
    const uint64_t o1 = stream_off[(size_t)img * n_streams + k + 1];
    uint64_t *so = suboff + (size_t)img * total_sub + d.sub_first;
    uint32_t *sl = sublen + (size_t)img * total_sub + d.sub_first;
    // the offsets come from the caller: they must start at 0, never decrease and end inside the blob, or every
    // substream of this stream reads as empty (zeros past the end, like torchac) and the call reports LLICTI_E_STREAM
    const uint64_t total = stream_off[(size_t)n_images * n_streams];
    if (!(o0 <= o1 && o1 <= total && total <= blob_bytes && o1 - o0 <= 0xFFFFFFFFull) || stream_off[0] != 0) {
        if (threadIdx.x == 0) atomicExch(status, LLICTI_E_STREAM);
        for (int j = threadIdx.x; j < d.S; j += blockDim.x) { so[j] = 0; sl[j] = 0; }
        return;
    }
    if (!sub_mode) {
        if (threadIdx.x == 0) { so[0] = o0; sl[0] = (uint32_t)(o1 - o0); }
        return;
    }
    const uint8_t *src = blob + o0;
    const uint64_t size = o1 - o0;
    const uint32_t hdr = 2 + 2 * d.S;
    if (size < hdr || (int)(src[0] | (src[1] << 8)) != d.S) {
        if (threadIdx.x == 0) atomicExch(status, LLICTI_E_STREAM);
        for (int j = threadIdx.x; j < d.S; j += blockDim.x) { so[j] = o0; sl[j] = 0; }
        return;
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ uint32_t wt[8];
    uint64_t running = hdr;
    for (int base = 0; base < d.S; base += 256) {
        const int j = base + threadIdx.x;
        const uint32_t v = j < d.S ? (uint32_t)(src[2 + 2 * j] | (src[3 + 2 * j] << 8)) : 0;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) wt[wid] = x;
        __syncthreads();
        uint32_t wp = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < wid) wp += wt[w];
            tot += wt[w];
        }
        if (j < d.S) {
            const uint64_t off = running + wp + x - v;
            so[j] = o0 + off;
            sl[j] = (off + v <= size) ? v : 0;
            if (off + v > size) atomicExch(status, LLICTI_E_STREAM);
        }
        running += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0 && running != size) atomicExch(status, LLICTI_E_STREAM);
}

// torchac.decode_int16_normalized_cdf from a dense table (parity entry point).
__global__ void __launch_bounds__(128)
decode_table_kernel(const int16_t *__restrict__ table, int n_sym, int Lp, int S, const uint8_t *__restrict__ in,
                    const uint32_t *__restrict__ offs, int16_t *__restrict__ sym) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= S) return;
    AcDecoder dec;
    dec.init(in + offs[j], offs[j + 1] - offs[j]);
    for (int i = j; i < n_sym; i += S) {
        const uint16_t *row = reinterpret_cast<const uint16_t *>(table) + (size_t)i * Lp;
        const uint32_t target = dec.target();
        int lo = 0, hi = Lp - 1;
        while (lo + 1 < hi) {
            const int mid = (lo + hi) >> 1;
            if (row[mid] <= target) lo = mid; else hi = mid;
        }
        sym[i] = (int16_t)lo;
        if (i + S < n_sym) dec.consume(row[lo], lo == Lp - 2 ? 0x10000u : (uint32_t)row[lo + 1]);
    }
}

// ------------------------------------------------------------------------------------------
// Self-test of the hoisted division (gmm.cuh: fdiv_hoisted) against div.rn.f32 on random operands
// drawn from the domain the CDF stage produces: sampling points of every alphabet, means near and
// far from them, spreads from the clamp value to 2^20 including hard mantissas.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
selftest_fdiv_kernel(uint64_t seed, int per_thread, NumericsProfile np, unsigned long long *mismatches) {
    uint64_t s = seed + 0x1234567ull * (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x);
    unsigned long long bad = 0;
    for (int it = 0; it < per_thread; ++it) {
        const uint64_t r0 = splitmix64(s), r1 = splitmix64(s);
        // spread: log-uniform over [2^-12, 2^20] with a random mantissa; every 8th one has a hard mantissa
        const int e = (int)(r0 % 33) - 12;
        uint32_t mant = (uint32_t)(r0 >> 8) & 0x7FFFFFu;
        if (((r0 >> 40) & 7) == 0) mant = ((r0 >> 43) & 1) ? 0x7FFFFFu - (uint32_t)((r0 >> 44) & 3) : (uint32_t)((r0 >> 44) & 3);
        float sigma = __uint_as_float(((uint32_t)(e + 127) << 23) | mant);
        sigma = fmaxf(sigma, (float)(0.11 / 255.0));
        // sampling point of a random alphabet, mean: near the point (a few ulps), image-range, or far
        const int min_val = (int)(r1 % 511) - 255, k = (int)((r1 >> 16) % 512);
        const CdfGrid g = make_grid(min_val, min_val + 510);
        const float pt = grid_point(g, k, np);
        const uint32_t sel = (uint32_t)(r1 >> 32) & 3u;
        float mu;
        if (sel == 0) mu = __uint_as_float(__float_as_uint(pt) + (uint32_t)((r1 >> 34) & 15) - 8u);
        else if (sel == 1) mu = (float)((r1 >> 34) & 0xFFFFF) * (1.5f / 1048576.0f) - 0.25f;
        else if (sel == 2) mu = ((float)((r1 >> 34) & 0xFFFFF) - 524288.0f) * (1.0f / 512.0f);
        else mu = pt - sigma * ((float)((r1 >> 34) & 0xFFFF) * (1.0f / 4096.0f) - 8.0f);
        GmmChannel c;
#pragma unroll
        for (int m = 0; m < kM; ++m) { c.sigma[m] = sigma; c.mu[m] = mu; c.w[m] = 0.2f; }
        gmm_prepare(c, np);
        if (!c.fast) continue;
        const float x = __fsub_rn(pt, c.mu[0]);
        const float a = fdiv_hoisted(x, c.sigma[0], c.rinv[0], 1), b = __fdiv_rn(x, c.sigma[0]);
        bad += __float_as_uint(a) != __float_as_uint(b);
        // the weight normalisation of gmm_prepare: five weights in [1e-6, 2^9] with random mantissas (every 8th hard), against
        // div.rn.f32 by the same denominator
        {
            const uint64_t r2 = splitmix64(s), r3 = splitmix64(s);
            GmmChannel cw;
            float wraw[kM];
#pragma unroll
            for (int m = 0; m < kM; ++m) {
                const uint64_t rr = m < 3 ? (r2 >> (21 * m)) : (r3 >> (21 * (m - 3)));
                const int ew = (int)(rr % 29) - 20;                                   // 2^-20 .. 2^8
                uint32_t mw = (uint32_t)(rr >> 5) & 0x7FFFFFu;
                if (((rr >> 3) & 7) == 0) mw = (rr & 4) ? 0x7FFFFFu - (uint32_t)(rr & 3) : (uint32_t)(rr & 3);
                wraw[m] = fmaxf(__uint_as_float(((uint32_t)(ew + 127) << 23) | mw), 1e-6f);
                cw.w[m] = wraw[m]; cw.sigma[m] = 0.01f; cw.mu[m] = 0.f;
            }
            const float den = __fadd_rn(sum5(wraw, np), 1e-9f);
            gmm_prepare(cw, np);
#pragma unroll
            for (int m = 0; m < kM; ++m) bad += __float_as_uint(cw.w[m]) != __float_as_uint(__fdiv_rn(wraw[m], den));
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

int launch_selftest_fdiv(llicti_ctx *ctx, long long n_pairs, uint64_t seed, unsigned long long *mismatches_dev, cudaStream_t st) {
    const int threads = 256, blocks = 1184, per_thread = (int)((n_pairs + (long long)threads * blocks - 1) / ((long long)threads * blocks));
    selftest_fdiv_kernel<<<blocks, threads, 0, st>>>(seed, per_thread, ctx->num, mismatches_dev);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

// ------------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------------
int launch_cdf_table(llicti_ctx *ctx, const float *params, const int16_t *yband, int clr, int min_val, int max_val,
                     int P, int16_t *table, cudaStream_t st) {
    const int warps_per_block = 4;
    cdf_table_kernel<<<(P + warps_per_block - 1) / warps_per_block, 128, 0, st>>>(params, yband, clr, min_val, max_val,
                                                                                 P, ctx->num, table);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_cdf_bounds_flat(llicti_ctx *ctx, const float *params, const int16_t *yband, int clr, int min_val,
                           int max_val, int P, uint32_t *bounds, cudaStream_t st) {
    cdf_bounds_flat_kernel<<<(P + 127) / 128, 128, 0, st>>>(params, yband, clr, min_val, max_val, P, ctx->num, bounds);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

static int stream_index(const Plan &p, int scale, int band, int clr) {
    return (p.g.num_scales - 1 - scale) * 9 + 3 * band + clr;
}

int launch_band_bounds(llicti_ctx *ctx, const Plan &p, int scale, int band, const float *params,
                       const int16_t *planes, const int32_t *minmax, int n, uint32_t *bounds, int64_t sym_stride,
                       cudaStream_t st) {
    ProfScope prof_(ctx, KC_BOUNDS, st);
    BandGeom bg;
    const StreamDesc &d0 = p.sd[stream_index(p, scale, band, 0)];
    bg.Hs = d0.Hs; bg.Ws = d0.Ws; bg.crop_h = d0.crop_h; bg.crop_w = d0.crop_w; bg.band = band;
    for (int c = 0; c < 3; ++c) bg.sym_off[c] = p.sd[stream_index(p, scale, band, c)].sym_off;
    dim3 grid((d0.n_sym + 127) / 128, n);
    band_bounds_kernel<<<grid, 128, 0, st>>>(params, planes, minmax, bg, ctx->num, bounds, sym_stride);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_encode_all(llicti_ctx *ctx, const Plan &p, const uint32_t *bounds, int64_t sym_stride, int n,
                      uint8_t *scratch, int64_t scratch_stride, uint32_t *sublen, cudaStream_t st) {
    ProfScope prof_(ctx, KC_ENCODE, st);
    const int total_sub = (int)p.g.substreams;
    if ((long long)total_sub * n <= kWarpChainLimit) {      // few, long chains: one warp each
        if (getenv("LLICTI_ENC_SINGLE_WARP")) {                 // A/B: one warp does both phases
            dim3 grid((total_sub + 3) / 4, n);
            encode_all_warp_kernel<<<grid, 128, 0, st>>>(ctx->d_sd, p.n_streams, total_sub, bounds, sym_stride, scratch,
                                                         scratch_stride, sublen, ctx->d_status);
        } else {
            dim3 grid((total_sub + 1) / 2, n);
            encode_all_pair_kernel<<<grid, 128, 0, st>>>(ctx->d_sd, p.n_streams, total_sub, bounds, sym_stride, scratch,
                                                         scratch_stride, sublen, ctx->d_status);
        }
    } else {
        dim3 grid((total_sub + 127) / 128, n);
        encode_all_kernel<<<grid, 128, 0, st>>>(ctx->d_sd, p.n_streams, total_sub, bounds, sym_stride, scratch,
                                                scratch_stride, sublen, ctx->d_status);
    }
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_encode_flat(llicti_ctx *ctx, const uint32_t *bounds, int n_sym, int S, uint8_t *out, int slot_bytes,
                       uint32_t *lens, cudaStream_t st) {
    if (S <= kWarpChainLimit)
        encode_flat_warp_kernel<<<(S + 3) / 4, 128, 0, st>>>(bounds, n_sym, S, out, slot_bytes, lens, ctx->d_status);
    else
        encode_flat_kernel<<<(S + 127) / 128, 128, 0, st>>>(bounds, n_sym, S, out, slot_bytes, lens, ctx->d_status);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_compact(llicti_ctx *ctx, const Plan &p, int n, const uint8_t *scratch, int64_t scratch_stride,
                   const uint32_t *sublen, uint64_t *stream_bytes, uint64_t *stream_off, uint8_t *out, size_t out_cap,
                   cudaStream_t st) {
    ProfScope prof_(ctx, KC_COMPACT, st);
    const int total_sub = (int)p.g.substreams;
    const int sub_mode = ctx->cfg.sub_len > 0;
    dim3 grid(p.n_streams, n);
    stream_sizes_kernel<<<grid, 128, 0, st>>>(ctx->d_sd, p.n_streams, total_sub, sub_mode, sublen, stream_bytes);
    scan_kernel<<<1, 1024, 0, st>>>(stream_bytes, stream_off, n * p.n_streams);
    gather_kernel<<<grid, 256, 0, st>>>(ctx->d_sd, p.n_streams, total_sub, sub_mode, scratch, scratch_stride, sublen,
                                        stream_off, out, (uint64_t)out_cap, ctx->d_status);
    ctx->launches += 3;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_index_streams(llicti_ctx *ctx, const Plan &p, int n, const uint8_t *blob, uint64_t blob_bytes, const uint64_t *stream_off,
                         uint64_t *suboff, uint32_t *sublen, cudaStream_t st) {
    ProfScope prof_(ctx, KC_INDEX, st);
    const int total_sub = (int)p.g.substreams;
    dim3 grid(p.n_streams, n);
    index_streams_kernel<<<grid, 256, 0, st>>>(ctx->d_sd, p.n_streams, total_sub, ctx->cfg.sub_len > 0, blob, blob_bytes, n,
                                               stream_off, suboff, sublen, ctx->d_status);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

int launch_decode_table(llicti_ctx *ctx, const int16_t *table, int n_sym, int Lp, int S, const uint8_t *in,
                        const uint32_t *offs, int16_t *sym, cudaStream_t st) {
    decode_table_kernel<<<(S + 127) / 128, 128, 0, st>>>(table, n_sym, Lp, S, in, offs, sym);
    ctx->launches += 1;
    LLICTI_CUDA(cudaGetLastError());
    return LLICTI_OK;
}

}  // namespace llicti
