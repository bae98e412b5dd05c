"""The closed forms the CUDA coder uses (llicti_b200/csrc/kernels_decode.cu, kernels_coder.cu), checked on the CPU
against torchac's bit-serial definition (as restated in oracle/llicti_oracle.py::ac_encode_table_py):

  * interval update in 32-bit arithmetic: (span * c) >> 16 == high word of span * (c << 16) for span < 2^32, the
    full range gives c << 16, an upper bound of 2^16 gives span;
  * renormalisation as ONE shift: sh = clz((nl ^ nh) & ~((nl & ~nh) << 1)) equals the number of E1/E2/E3 iterations,
    and the shifted registers / pending-bit count / flipped value bit equal the loop's;
  * clz through the exponent of a round-toward-zero float conversion;
  * the decoder's search key: low + ((span * q) >> 16) <= value  <=>  q <= ((value - low + 1) * 2^16 - 1) // span.
"""
import random

MASK = 0xFFFFFFFF


def serial_update(low, high, value, c_low, c_high, bits):
    """torchac's loop on (low, high) and the decoder's `value`; returns (low, high, value, shifts, underflows)."""
    span = high - low + 1
    high = low - 1 + ((span * c_high) >> 16)
    low = low + ((span * c_low) >> 16)
    shifts = under = 0
    while True:
        if high < 0x80000000 or low >= 0x80000000:
            low = (low << 1) & MASK
            high = ((high << 1) & MASK) | 1
            value = ((value << 1) & MASK) | bits.pop(0)
        elif low >= 0x40000000 and high < 0xC0000000:
            under += 1
            low = (low << 1) & 0x7FFFFFFF
            high = ((high << 1) & MASK) | 0x80000001
            value = (((value - 0x40000000) << 1) & MASK) | bits.pop(0)
        else:
            break
        shifts += 1
    return low, high, value, shifts, under


def clz32(x):
    return 32 - x.bit_length()


def closed_update(low, high, value, c_low, c_high, next_bits):
    span32 = (high - low + 1) & MASK                      # 0 = 2^32
    cl16, ch16 = (c_low << 16) & MASK, (c_high << 16) & MASK
    tl = cl16 if span32 == 0 else (span32 * cl16) >> 32
    th = span32 if ch16 == 0 else (ch16 if span32 == 0 else (span32 * ch16) >> 32)
    nl, nh = (low + tl) & MASK, (low + th - 1) & MASK
    sh = clz32((nl ^ nh) & ~(((nl & ~nh) << 1) & MASK) & MASK)
    ls = (nl << sh) & MASK
    new_low = ls & 0x7FFFFFFF
    new_high = ((nh << sh) & MASK) | ((1 << sh) - 1) | 0x80000000
    new_value = ((((value << 32) | next_bits) << sh) >> 32 & MASK) ^ (ls & 0x80000000)
    under = clz32((nl ^ nh) & ~(((nl & ~nh) << 1) & MASK) & MASK) - clz32(nl ^ nh)
    return new_low, new_high, new_value, sh, under, nl, nh


def random_state(rng):
    """A state satisfying the coder's invariant after renormalisation (or the initial full range)."""
    if rng.random() < 0.05:
        return 0, MASK
    while True:
        low, high = rng.randrange(0, 1 << 31), rng.randrange(1 << 31, 1 << 32)
        if low < 0x40000000 or high >= 0xC0000000:
            return low, high


def test_interval_update_and_single_shift_renormalisation_match_the_bit_serial_loop():
    rng = random.Random(1234)
    for _ in range(20000):
        low, high = random_state(rng)
        c_low = rng.randrange(0, 65536)
        c_high = rng.choice([65536, min(65536, c_low + 1 + int(rng.expovariate(1 / 300.0)))])
        value = rng.randrange(low, high + 1)
        # keep the symbol consistent with `value` for the decoder side of the identity
        span = high - low + 1
        nl_ref, nh_ref = low + ((span * c_low) >> 16), low - 1 + ((span * c_high) >> 16)
        if not (nl_ref <= nh_ref):
            continue
        value = rng.randrange(nl_ref, nh_ref + 1)
        stream = [rng.getrandbits(1) for _ in range(64)]
        next_bits = int("".join(map(str, stream[:32])), 2)
        r_low, r_high, r_value, r_sh, r_under = serial_update(low, high, value, c_low, c_high, list(stream))
        n_low, n_high, n_value, sh, under, nl, nh = closed_update(low, high, value, c_low, c_high, next_bits)
        assert (nl, nh) == (nl_ref, nh_ref)
        assert (n_low, n_high, sh, under) == (r_low, r_high, r_sh, r_under)
        assert n_value == r_value
        assert sh <= 31


def test_clz_through_round_toward_zero_float_exponent():
    import numpy as np
    rng = random.Random(7)
    xs = [1, 2, 3, (1 << 24) - 1, 1 << 24, (1 << 25) - 1, (1 << 31) - 1, 1 << 31, MASK] + [rng.randrange(1, 1 << 32) for _ in range(5000)]
    bumped = 0
    for x in xs:
        keep = max(x.bit_length() - 24, 0)
        f_rz = np.float32((x >> keep) << keep)                       # what I2F.RZ yields: the mantissa is truncated
        assert float(f_rz) == float((x >> keep) << keep)
        e = int(f_rz.view(np.uint32)) >> 23
        assert 158 - e == clz32(x)
        bumped += (158 - (int(np.float32(x).view(np.uint32)) >> 23)) != clz32(x)
    assert bumped > 0      # round-to-nearest would be wrong just below a power of two: the kernels must use .rz


def test_decoder_search_key_needs_no_division():
    rng = random.Random(99)
    for _ in range(20000):
        low, high = random_state(rng)
        span = high - low + 1
        value = rng.randrange(low, high + 1)
        q = rng.randrange(0, 65537)
        key = ((value - low + 1) * 65536 - 1) // span
        assert (low + ((span * q) >> 16) <= value) == (q <= key)


def _torchac_search(cdf, count):
    """Largest m in [0, Lp-2] with cdf[m] <= count (torchac's binary search; cdf[Lp-1] acts as 2^16)."""
    left, right = 0, len(cdf) - 1
    while left + 1 < right:
        m = (left + right) // 2
        if cdf[m] <= count:
            left = m
        else:
            right = m
    return left


def test_window_selection_rule_finds_torchacs_symbol_or_flags_the_chunk():
    """Python model of decode_step_fast's lane logic: slots q(base .. base+31) with q(last) = 2^16 stored as 0, the
    candidate lanes 0..30 (fewer for small alphabets), ch16 == 0 meaning an upper bound of 2^16, `bad` for a symbol
    outside the window or a full-range state."""
    rng = random.Random(4321)
    checked = flagged = 0
    for _ in range(4000):
        Lp = rng.choice([2, 5, 17, 32, 33, 64, 257, 400, 512])
        last = Lp - 1
        steps = sorted(rng.sample(range(1, 65536), last - 1)) if last > 1 else []
        cdf = [0] + steps + [65536]                         # strictly increasing, cdf[last] = 2^16
        low, high = random_state(rng)
        span = high - low + 1
        value = rng.randrange(low, high + 1)
        count = ((value - low + 1) * 65536 - 1) // span
        sym = _torchac_search(cdf, min(count, 65535))
        kc = rng.randrange(0, last) if last > 0 else 0
        base = min(max(kc - 15, 0), max(last - 31, 0))
        slots = [(cdf[base + l] & 0xFFFF) if base + l <= last else 0 for l in range(32)]
        vmask = 0x7FFFFFFF if last >= 31 else (1 << last) - 1
        span32 = span & MASK
        ballot = 0
        cand = []
        for l in range(32):
            cl16 = slots[l] << 16
            ch16 = (slots[l + 1] << 16) if l < 31 else cl16   # shfl_down: lane 31 keeps its own value
            nl = (low + ((span32 * cl16) >> 32)) & MASK
            nhp1 = (low + (span32 if ch16 == 0 else (span32 * ch16) >> 32)) & MASK
            if value >= nl:
                ballot |= 1 << l
            cand.append((nl, nhp1))
        li = bin(ballot & vmask).count("1") - 1
        bad = span32 == 0 or li < 0 or value >= cand[li & 31][1]
        inside = base <= sym <= base + 30 and sym < last
        if span32 != 0 and inside:
            assert not bad and base + li == sym, (Lp, base, sym, li)
            checked += 1
        else:
            assert bad or base + li == sym
            flagged += bad
    assert checked > 1000 and flagged > 100


def test_search_stand_in_of_the_normal_cdf_is_accurate_to_1e6():
    """The lane decoder's search (kernels_decode.cu: approx_q_slope) evaluates the normal CDF through Abramowitz & Stegun
    7.1.26 in fp32 with approximate reciprocal and exp2.  Restated here in fp32 numpy against scipy's erfc: the error stays
    below 1e-6 of the CDF, i.e. below a tenth of a table unit (the logistic stand-in it replaced: 3e-4, 20 units), and the
    density it gets from the same exponential matches the derivative."""
    import numpy as np
    from scipy.special import erfc
    z = np.linspace(-9, 9, 200001).astype(np.float32)
    x = np.abs(z) * np.float32(0.70710678)
    e = np.exp2(x * x * np.float32(-1.4426950)).astype(np.float32)
    t = (np.float32(1) / (np.float32(0.3275911) * x + np.float32(1))).astype(np.float32)
    poly = t * (t * (t * (t * (t * np.float32(1.061405429) + np.float32(-1.453152027)) + np.float32(1.421413741)) + np.float32(-0.284496736)) + np.float32(0.254829592))
    h = np.float32(0.5) * poly * e
    cdf = np.where(z >= 0, np.float32(1) - h, h).astype(np.float64)
    ref = 0.5 * erfc(-z.astype(np.float64) / np.sqrt(2.0))
    assert np.max(np.abs(cdf - ref)) < 1e-6
    dens = 0.39894228 * e.astype(np.float64)
    assert np.max(np.abs(dens - np.exp(-0.5 * z.astype(np.float64) ** 2) / np.sqrt(2 * np.pi))) < 1e-6
