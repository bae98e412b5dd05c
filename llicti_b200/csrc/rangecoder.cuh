// Device-side binary arithmetic coder, bit-compatible with torchac 0.9.3 (the coder the
// reference calls at graphs/models/LLICTI_nets.py:406-407 and :492-493; algorithm restated in
// oracle/torchac_port.c).  One coder state per thread; a torchac stream is S = 1, the
// throughput container runs S interleaved substreams with the very same arithmetic.
//
// The bit-serial renormalisation loop of the original (one E1/E2/E3 decision per output bit)
// is collapsed into two count-leading-zeros steps that produce the identical bit sequence:
//   n = clz(low ^ high)  leading equal bits are shifted out together (E1/E2 run),
//   k = leading run where low has 1s and high has 0s below the MSB (E3 / underflow run).
#pragma once

#include "common.cuh"

namespace llicti {

// MSB-first bit sink writing big-endian 32-bit words into a 4-byte aligned slot.
struct BitWriter {
    uint8_t *base;
    uint32_t cap;      // slot capacity in bytes
    uint32_t nbytes;   // bytes written so far (multiple of 4 until finish)
    uint64_t acc;      // pending bits, right aligned
    int nbits;         // number of pending bits (< 32 between calls)
    int overflow;
    bool writer;       // false in the lanes of a warp-per-chain encoder that only mirror the state

    __device__ __forceinline__ void init(uint8_t *b, uint32_t c, bool wr = true) {
        base = b; cap = c; nbytes = 0; acc = 0; nbits = 0; overflow = 0; writer = wr;
    }
    __device__ __forceinline__ void put(uint32_t bits, int count) {   // 0 <= count <= 32
        acc = (acc << count) | bits;
        nbits += count;
        if (nbits >= 32) {
            const uint32_t word = (uint32_t)(acc >> (nbits - 32));
            if (nbytes + 4 <= cap) { if (writer) *reinterpret_cast<uint32_t *>(base + nbytes) = __byte_perm(word, 0, 0x0123); }
            else overflow = 1;
            nbytes += 4;
            nbits -= 32;
            acc &= (1ull << nbits) - 1ull;
        }
    }
    __device__ __forceinline__ void put_run(uint32_t bit, uint32_t count) {   // `count` copies of `bit`
        const uint32_t pat = bit ? 0xFFFFFFFFu : 0u;
        while (count >= 32) { put(pat, 32); count -= 32; }
        if (count) put(pat >> (32 - count), (int)count);
    }
    __device__ __forceinline__ uint32_t finish() {   // zero-pad to a byte boundary
        const int nb = (nbits + 7) >> 3;
        const uint32_t word = nbits ? (uint32_t)(acc << (32 - nbits)) : 0u;
        for (int i = 0; i < nb; ++i) {
            if (nbytes + i < cap) { if (writer) base[nbytes + i] = (uint8_t)(word >> (24 - 8 * i)); }
            else overflow = 1;
        }
        nbytes += nb;
        nbits = 0;
        return nbytes;
    }
};

struct AcEncoder {
    uint32_t low, high, pending;
    BitWriter bw;

    __device__ __forceinline__ void init(uint8_t *slot, uint32_t cap, bool writer = true) {
        low = 0; high = 0xFFFFFFFFu; pending = 0;
        bw.init(slot, cap, writer);
    }
    // Interval update + renormalisation.  n = equal leading bits (emitted), k = underflow run
    // (counted in `pending`); d & ~(m << 1) has its first one at n + k (see next_state() in
    // kernels_decode.cu), so the registers are shifted once, branch-free; only the bit output
    // of a step that resolves pending underflow bits takes a branch.
    __device__ __forceinline__ void encode(uint32_t c_low, uint32_t c_high) {
        const uint32_t sm1 = high - low;                               // span - 1; span * c = sm1 * c + c
        const uint32_t nl = low + (uint32_t)(((uint64_t)sm1 * c_low + c_low) >> 16);
        const uint32_t nh = (low - 1u) + (uint32_t)(((uint64_t)sm1 * c_high + c_high) >> 16);
        const uint32_t d = nl ^ nh;
        const int n = __clz(d);
        const int sh = __clz(d & ~((nl & ~nh) << 1));
        if (__builtin_expect(pending != 0u && n > 0, 0)) {
            const uint32_t b = nl >> 31;
            bw.put(b, 1);
            bw.put_run(b ^ 1u, pending);
            pending = 0;
            if (n > 1) bw.put((nl << 1) >> (33 - n), n - 1);
        } else {
            bw.put(__funnelshift_l(nl, 0u, n), n);                     // the n leading bits of nl (n = 0: nothing)
        }
        pending += (uint32_t)(sh - n);
        low = (nl << sh) & 0x7FFFFFFFu;
        high = (nh << sh) | ~(0xFFFFFFFFu << sh) | 0x80000000u;
    }
    __device__ __forceinline__ uint32_t finish() {
        pending += 1;
        const uint32_t b = low < 0x40000000u ? 0u : 1u;
        bw.put(b, 1);
        bw.put_run(b ^ 1u, pending);
        return bw.finish();
    }
};

// MSB-first bit source over [p, p + len); reads zeros past the end like torchac.
struct BitReader {
    const uint8_t *p;
    uint32_t pos, len;
    uint64_t buf;   // upcoming bits, left aligned
    int avail;

    __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t n) {
        p = ptr; pos = 0; len = n; buf = 0; avail = 0;
    }
    __device__ __forceinline__ uint32_t take(int n) {   // 1 <= n <= 32
        if (avail < n) {
            while (avail <= 56 && pos < len) {
                buf |= (uint64_t)p[pos++] << (56 - avail);
                avail += 8;
            }
        }
        const uint32_t v = (uint32_t)(buf >> (64 - n));
        buf <<= n;
        avail = max(avail - n, 0);
        return v;
    }
};

struct AcDecoder {
    uint32_t low, high, value;
    BitReader br;

    __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t n) {
        low = 0; high = 0xFFFFFFFFu;
        br.init(ptr, n);
        value = br.take(32);
    }
    // 16-bit cumulative count the next symbol must bracket.
    __device__ __forceinline__ uint32_t target() const {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1ull;
        const uint64_t num = (((uint64_t)value - (uint64_t)low + 1ull) << 16) - 1ull;
        return (uint32_t)(num / span) & 0xFFFFu;
    }
    __device__ __forceinline__ void consume(uint32_t c_low, uint32_t c_high) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1ull;
        high = (low - 1u) + (uint32_t)((span * c_high) >> 16);
        low = low + (uint32_t)((span * c_low) >> 16);
        const int n = __clz(low ^ high);
        if (n > 0) {
            low <<= n;
            high = (high << n) | ((1u << n) - 1u);
            value = (value << n) | br.take(n);
        }
        const uint32_t y = (low << 1) & ~(high << 1);
        const int k = __clz(~y);
        if (k > 0) {
            low = (low << k) & 0x7FFFFFFFu;
            high = (high << k) | 0x80000000u | ((1u << k) - 1u);
            value = ((value << k) | br.take(k)) ^ 0x80000000u;
        }
    }
};

// MSB-first bit source reading aligned 32-bit words two words ahead of the coder, so no memory
// latency sits on the serial decode chain.  Bytes outside [ptr, ptr + n) read as zero (torchac
// reads zeros past the end of the stream).
struct BitReaderW {
    const uint32_t *w;        // 4-byte aligned base (<= ptr)
    uint32_t lo_byte, hi_byte;   // valid bytes [lo, hi) relative to w
    uint32_t idx;             // next word to fetch
    uint32_t a0, a1;          // the two upcoming words
    uint64_t buf;             // upcoming bits, left aligned
    int avail;

    __device__ __forceinline__ uint32_t fetch(uint32_t i) const {
        const uint32_t b0 = i * 4u;
        if (b0 >= hi_byte) return 0u;
        uint32_t v = __byte_perm(__ldg(w + i), 0, 0x0123);   // first byte in the most significant position
        if (b0 < lo_byte) v &= 0xFFFFFFFFu >> (8u * (lo_byte - b0));
        if (b0 + 4u > hi_byte) v &= 0xFFFFFFFFu << (8u * (b0 + 4u - hi_byte));
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t n) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(ptr);
        w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
        lo_byte = (uint32_t)(a & 3);
        hi_byte = lo_byte + n;
        const uint32_t w0 = fetch(0);
        buf = (uint64_t)w0 << (32 + 8 * lo_byte);
        avail = 32 - 8 * (int)lo_byte;
        a0 = fetch(1);
        a1 = fetch(2);
        idx = 3;
    }
    __device__ __forceinline__ uint32_t take(int n) {   // 1 <= n <= 32
        if (avail < n) {
            buf |= (uint64_t)a0 << (32 - avail);
            avail += 32;
            a0 = a1;
            a1 = fetch(idx++);
        }
        const uint32_t v = (uint32_t)(buf >> (64 - n));
        buf <<= n;
        avail -= n;
        return v;
    }
};

// Same arithmetic as AcDecoder over the look-ahead bit source.
struct AcDecoderW {
    uint32_t low, high, value;
    BitReaderW br;

    __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t n) {
        low = 0; high = 0xFFFFFFFFu;
        br.init(ptr, n);
        value = br.take(32);
    }
    __device__ __forceinline__ void consume(uint32_t c_low, uint32_t c_high) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1ull;
        high = (low - 1u) + (uint32_t)((span * c_high) >> 16);
        low = low + (uint32_t)((span * c_low) >> 16);
        const int n = __clz(low ^ high);
        if (n > 0) {
            low <<= n;
            high = (high << n) | ((1u << n) - 1u);
            value = (value << n) | br.take(n);
        }
        const uint32_t y = (low << 1) & ~(high << 1);
        const int k = __clz(~y);
        if (k > 0) {
            low = (low << k) & 0x7FFFFFFFu;
            high = (high << k) | 0x80000000u | ((1u << k) - 1u);
            value = ((value << k) | br.take(k)) ^ 0x80000000u;
        }
    }
};

}  // namespace llicti
