"""oracle/torchac_port.c against its pure-Python twin and against itself (round trips)."""
import numpy as np
import pytest

from oracle import llicti_oracle as O


def random_tables(rng, n, Lp):
    """Strictly increasing uint16 rows built like the reference builds them
    (round(cdf * (65536 - (Lp-1))) + arange)."""
    pdf = rng.random((n, Lp - 1)) ** 4 + 1e-6
    cdf = np.concatenate([np.zeros((n, 1)), np.cumsum(pdf, axis=1)], axis=1)
    cdf /= cdf[:, -1:]
    q = np.rint(cdf * (65536 - (Lp - 1))).astype(np.int64) + np.arange(Lp)
    return (q & 0xFFFF).astype(np.uint16).view(np.int16)


@pytest.mark.parametrize("Lp", [2, 3, 17, 257, 512])
def test_c_matches_python_twin(Lp):
    rng = np.random.default_rng(Lp)
    n = 400
    tab = random_tables(rng, n, Lp)
    sym = rng.integers(0, Lp - 1, size=n).astype(np.int16)
    assert O.ac_encode_table(tab, sym) == O.ac_encode_table_py(tab, sym)


@pytest.mark.parametrize("Lp,n", [(2, 50), (257, 5000), (512, 3000), (40, 1)])
def test_round_trip(Lp, n):
    rng = np.random.default_rng(n + Lp)
    tab = random_tables(rng, n, Lp)
    sym = rng.integers(0, Lp - 1, size=n).astype(np.int16)
    blob = O.ac_encode_table(tab, sym)
    assert np.array_equal(O.ac_decode_table(tab, blob), sym)


def test_bounds_entry_equals_table_entry():
    rng = np.random.default_rng(3)
    n, Lp = 3000, 300
    tab = random_tables(rng, n, Lp)
    sym = rng.integers(0, Lp - 1, size=n).astype(np.int16)
    u = tab.view(np.uint16).astype(np.uint32)
    lo = u[np.arange(n), sym]
    hi = np.where(sym == Lp - 2, 0x10000, u[np.arange(n), np.minimum(sym + 1, Lp - 1)])
    bounds = (lo | ((hi - 1) << 16)).astype(np.uint32)
    assert O.ac_encode_bounds(bounds) == O.ac_encode_table(tab, sym)


def test_skewed_symbols_and_max_symbol():
    # every symbol the max symbol (c_high = 0x10000) and every symbol the rarest one
    rng = np.random.default_rng(4)
    n, Lp = 2000, 257
    tab = random_tables(rng, n, Lp)
    for sym in (np.full(n, Lp - 2, dtype=np.int16), np.zeros(n, dtype=np.int16)):
        blob = O.ac_encode_table(tab, sym)
        assert np.array_equal(O.ac_decode_table(tab, blob), sym)


def test_substream_container_round_trip():
    parts = [b"", b"abc", bytes(range(200))]
    blob = O.pack_substreams(parts)
    assert blob[:2] == (3).to_bytes(2, "little")
    assert O.unpack_substreams(blob) == parts
    assert O.num_substreams(10, 0 + 4096) == 1
    assert O.num_substreams(98304, 2048) == 64      # rounded up to a multiple of 32 past 32
    assert O.num_substreams(40000, 2048) == 20
