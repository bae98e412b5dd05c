/*
 * oracle/torchac_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic coder the reference calls at
 *   graphs/models/LLICTI_nets.py:406-407  torchac.encode_int16_normalized_cdf(cdf, sym)
 *   graphs/models/LLICTI_nets.py:492-493  torchac.decode_int16_normalized_cdf(cdf, bytes)
 *
 * The coder itself lives in the third-party wheel torchac==0.9.3 (README.md:16 of
 * the reference), which is NOT vendored under /root/reference and is not
 * installable here (no network).  This file restates its published algorithm:
 * a 32-bit binary arithmetic coder (low/high registers, E1/E2/E3 renormalisation
 * with pending bits, MSB-first bit packing) working on 16-bit CDF rows that are
 * int16 in memory but read as uint16, where the right bound of the largest
 * symbol (Lp-2) is hard-wired to 0x10000 and the last table column is never
 * read.
 *
 * PARITY STATUS: "parity unpinned" with respect to the real torchac binary --
 * the reference ships no golden bitstreams and the wheel is absent.  What IS
 * pinned: the unmodified reference model code, driven through this coder,
 * round-trips losslessly (tests/golden/make_golden.py), and every CUDA coder
 * path is byte-identical to this restatement.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load this file.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

typedef struct {
    uint8_t *buf;
    size_t cap;
    size_t n;       /* bytes written (may exceed cap: then output is truncated) */
    uint8_t cache;
    uint8_t count;
} bitsink;

static void sink_put(bitsink *s, int bit) {
    s->cache = (uint8_t)((s->cache << 1) | (bit & 1));
    s->count++;
    if (s->count == 8) {
        if (s->n < s->cap) s->buf[s->n] = s->cache;
        s->n++;
        s->count = 0;
        s->cache = 0;
    }
}

static void sink_put_with_pending(bitsink *s, int bit, uint64_t *pending) {
    sink_put(s, bit);
    while (*pending > 0) {
        sink_put(s, !bit);
        (*pending)--;
    }
}

static void sink_flush(bitsink *s) {
    while (s->count != 0) sink_put(s, 0);
}

/* One coding step shared by the two encoder entry points. */
static void enc_step(bitsink *s, uint32_t *low, uint32_t *high, uint64_t *pending,
                     uint32_t c_low, uint32_t c_high) {
    const uint64_t span = (uint64_t)(*high) - (uint64_t)(*low) + 1;
    *high = (*low - 1u) + (uint32_t)((span * (uint64_t)c_high) >> 16);
    *low = (*low) + (uint32_t)((span * (uint64_t)c_low) >> 16);
    for (;;) {
        if (*high < 0x80000000u) {
            sink_put_with_pending(s, 0, pending);
            *low <<= 1;
            *high = (*high << 1) | 1u;
        } else if (*low >= 0x80000000u) {
            sink_put_with_pending(s, 1, pending);
            *low <<= 1;
            *high = (*high << 1) | 1u;
        } else if (*low >= 0x40000000u && *high < 0xC0000000u) {
            (*pending)++;
            *low = (*low << 1) & 0x7FFFFFFFu;
            *high = (*high << 1) | 0x80000001u;
        } else {
            break;
        }
    }
}

static size_t enc_finish(bitsink *s, uint32_t low, uint64_t pending) {
    pending += 1;
    sink_put_with_pending(s, low < 0x40000000u ? 0 : 1, &pending);
    sink_flush(s);
    return s->n;
}

/*
 * Encode n symbols against a dense table cdf[n][Lp] (int16 storage, uint16
 * meaning).  Returns the number of bytes the stream needs; at most `cap` of
 * them are stored in `out`.  Symbols must lie in [0, Lp-2].
 */
EXPORT size_t oracle_ac_encode_table(const int16_t *cdf, const int16_t *sym, size_t n, int Lp,
                                     uint8_t *out, size_t cap) {
    bitsink s = {out, cap, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;
    const int max_symbol = Lp - 2;
    for (size_t i = 0; i < n; ++i) {
        const uint16_t *row = (const uint16_t *)(cdf + i * (size_t)Lp);
        const int v = sym[i];
        const uint32_t c_low = row[v];
        const uint32_t c_high = (v == max_symbol) ? 0x10000u : row[v + 1];
        enc_step(&s, &low, &high, &pending, c_low, c_high);
    }
    return enc_finish(&s, low, pending);
}

/*
 * Same coder fed with per-symbol bounds instead of a table: bounds[i] packs
 * c_low in the low 16 bits and (c_high - 1) in the high 16 bits.
 */
EXPORT size_t oracle_ac_encode_bounds(const uint32_t *bounds, size_t n, uint8_t *out, size_t cap) {
    bitsink s = {out, cap, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu;
    uint64_t pending = 0;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t c_low = bounds[i] & 0xFFFFu;
        const uint32_t c_high = (bounds[i] >> 16) + 1u;
        enc_step(&s, &low, &high, &pending, c_low, c_high);
    }
    return enc_finish(&s, low, pending);
}

typedef struct {
    const uint8_t *buf;
    size_t len;
    size_t pos;
    uint8_t cache;
    uint8_t cached_bits;
} bitsource;

static void source_get(bitsource *s, uint32_t *value) {
    if (s->cached_bits == 0) {
        if (s->pos == s->len) {
            *value <<= 1; /* zeros past the end */
            return;
        }
        s->cache = s->buf[s->pos++];
        s->cached_bits = 8;
    }
    *value = (*value << 1) | ((uint32_t)(s->cache >> (s->cached_bits - 1)) & 1u);
    s->cached_bits--;
}

static uint16_t table_search(const uint16_t *row, uint16_t target, uint16_t max_sym) {
    uint16_t left = 0;
    uint16_t right = (uint16_t)(max_sym + 1);
    while (left + 1 < right) {
        const uint16_t m = (uint16_t)((left + right) / 2);
        const uint16_t v = row[m];
        if (v < target) left = m;
        else if (v > target) right = m;
        else return m;
    }
    return left;
}

/* Decode n symbols against a dense table cdf[n][Lp]. */
EXPORT void oracle_ac_decode_table(const int16_t *cdf, size_t n, int Lp, const uint8_t *in, size_t in_len,
                                   int16_t *sym_out) {
    bitsource src = {in, in_len, 0, 0, 0};
    uint32_t low = 0, high = 0xFFFFFFFFu, value = 0;
    const int max_symbol = Lp - 2;
    for (int i = 0; i < 32; ++i) source_get(&src, &value);
    for (size_t i = 0; i < n; ++i) {
        const uint64_t span = (uint64_t)high - (uint64_t)low + 1;
        const uint16_t count =
            (uint16_t)((((uint64_t)value - (uint64_t)low + 1) * 0x10000ull - 1) / span);
        const uint16_t *row = (const uint16_t *)(cdf + i * (size_t)Lp);
        const uint16_t v = table_search(row, count, (uint16_t)max_symbol);
        sym_out[i] = (int16_t)v;
        if (i == n - 1) break;
        const uint32_t c_low = row[v];
        const uint32_t c_high = (v == max_symbol) ? 0x10000u : row[v + 1];
        high = (low - 1u) + (uint32_t)((span * (uint64_t)c_high) >> 16);
        low = low + (uint32_t)((span * (uint64_t)c_low) >> 16);
        for (;;) {
            if (low >= 0x80000000u || high < 0x80000000u) {
                low <<= 1;
                high = (high << 1) | 1u;
                source_get(&src, &value);
            } else if (low >= 0x40000000u && high < 0xC0000000u) {
                low = (low << 1) & 0x7FFFFFFFu;
                high = (high << 1) | 0x80000001u;
                value -= 0x40000000u;
                source_get(&src, &value);
            } else {
                break;
            }
        }
    }
}
