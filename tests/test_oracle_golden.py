"""The CPU oracle against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Integer stages and byte streams must match bit for bit; the
two float stages are compared bit-exactly when this host's vector erfc / reduction order is
the one the vectors were made with (canary), and by tolerance always."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import FORWARD_CASES, GOLDEN_CASES, TRAINED_CASES, load_golden, oracle_config_for, trained_state_dict
from oracle import llicti_oracle as O


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def host_matches_canary():
    c = load_golden("canary")
    x = torch.linspace(-6, 6, 4001, dtype=torch.float32)
    w = torch.from_numpy(np.random.default_rng(5).random((1, 37, 53, 1, 5), dtype=np.float32))
    wp = w.permute(0, 4, 1, 2, 3).contiguous().permute(0, 2, 3, 4, 1)
    return (digest(torch.erfc(x).numpy()) == str(c["erfc_digest"])
            and digest(torch.sum(wp, dim=4).numpy()) == str(c["sum_digest"]))


@pytest.fixture(scope="module")
def dumps():
    out = {}
    for name in GOLDEN_CASES:
        g = load_golden(name)
        cfg = oracle_config_for(name)
        sd = O.synthetic_state_dict(cfg)
        d = O.StageDump()
        bsl = O.OracleCodec(cfg, sd).compress(g["rgb"], d)
        out[name] = (g, cfg, sd, d, bsl)
    return out


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_integer_stages_bit_exact(dumps, name):
    g, cfg, sd, d, bsl = dumps[name]
    assert np.array_equal(d.ycocg, g["ycocg"])
    assert d.minmax == [int(v) for v in g["minmax"]]
    assert d.pad_int == int(g["pad_int"])
    assert np.array_equal(np.array(d.pad_flags, dtype=np.uint8), g["pad_flags"])
    for s in range(len(cfg.dwtlevels)):
        assert np.array_equal(d.planes[s], g[f"planes_{s}"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_header_bit_exact(dumps, name):
    g, cfg, sd, d, bsl = dumps[name]
    for j in range(9):
        assert bsl[0][j] == g[f"stream_0_{j}"].tobytes()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_network_outputs_close(dumps, name):
    g, cfg, sd, d, bsl = dumps[name]
    M = cfg.num_mixtures
    for (s, b), p in d.params.items():
        sub = p.reshape(12 * M, -1)[:, ::7]
        np.testing.assert_allclose(sub, g[f"params_{s}_{b}_sub"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_float_stages_and_streams_bit_exact_on_matching_host(dumps, name):
    if not host_matches_canary():
        pytest.skip("this host's torch erfc / reduction order differs from the one the golden vectors were made on")
    g, cfg, sd, d, bsl = dumps[name]
    for (s, b), p in d.params.items():
        assert digest(p) == str(g[f"params_{s}_{b}_digest"])
    for (s, b, c), t in d.tables.items():
        assert digest(t) == str(g[f"table_{s}_{b}_{c}_digest"])
        assert np.array_equal(t[::37], g[f"table_{s}_{b}_{c}_sub"])
    for i in range(1, len(bsl)):
        for j in range(9):
            assert bsl[i][j] == g[f"stream_{i}_{j}"].tobytes(), f"stream {i},{j}"
    assert sum(len(x) for r in bsl for x in r) == int(g["total_bytes"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_decodes_reference_streams(dumps, name):
    if not host_matches_canary():
        pytest.skip("needs the golden host's float behaviour")
    g, cfg, sd, d, bsl = dumps[name]
    S = len(cfg.dwtlevels)
    ref = [[g[f"stream_{i}_{j}"].tobytes() for j in range(9)] for i in range(S + 1)]
    assert np.array_equal(O.OracleCodec(cfg, sd).decompress(ref), g["rgb"])


@pytest.mark.parametrize("sub_len", [0, 100])
def test_oracle_round_trip_modes(sub_len):
    cfg = O.OracleConfig()
    sd = O.synthetic_state_dict(cfg)
    img = O.synthetic_image(40, 56, 9)
    codec = O.OracleCodec(cfg, sd, sub_len=sub_len)
    assert np.array_equal(codec.decompress(codec.compress(img)), img)


@pytest.mark.parametrize("threads", [1, 8, 16, 32, 64])
def test_oracle_round_trips_in_the_product_process_state(threads):
    """The oracle must reproduce itself (compress -> decompress) in a process that has loaded the product package
    and done torch CPU work at another thread count first (round 1: bench.py lost its cpu_baseline leg on one host
    to an oracle that did not), whatever the thread count; diagnose_round_trip names the stage if it does not."""
    import torch
    import llicti_b200  # noqa: F401  (the product package, as in bench.py's B200 arm)
    from oracle import llicti_oracle as O
    x = torch.randn(512, 512)
    (x @ x).sum().item()                                   # warms the intra-op pool at the default thread count
    before = torch.get_num_threads()
    try:
        torch.set_num_threads(threads)
        for cfg, (H, W) in ((O.OracleConfig(dwtlevels=(0, 1), chs=60), (64, 96)), (O.OracleConfig(), (53, 77))):
            codec = O.OracleCodec(cfg, O.synthetic_state_dict(cfg))
            img = O.synthetic_image(H, W, 11)
            rec = codec.decompress(codec.compress(img))
            assert np.array_equal(rec, img), O.diagnose_round_trip(codec, img)
            assert O.diagnose_round_trip(codec, img) == "this repetition round-tripped"
    finally:
        torch.set_num_threads(before)


@pytest.mark.parametrize("name", FORWARD_CASES)
def test_forward_self_informations_against_the_reference(name):
    """oracle.forward_self_informations restates LLICTI.forward (the rate-estimation path of validate / training);
    the fixtures are the unmodified reference's outputs (tests/golden/make_golden.py checks equality bit for bit on
    the host that makes them).  Bit-exact where this host's erfc / reduction order is the golden host's, and within
    a tight tolerance in bits everywhere."""
    g = load_golden(name)
    cfg = oracle_config_for(name)
    sd = O.synthetic_state_dict(cfg)
    mine = O.forward_self_informations(cfg, O.OracleNet(cfg, sd), g["rgb"])
    assert len(mine) == len(cfg.dwtlevels)
    exact = host_matches_canary()
    for s, q in enumerate(mine):
        ref = g[f"sinfo_{s}"]
        assert q.shape == ref.shape and q.shape[0] == 9
        if exact:
            assert np.array_equal(q, ref), f"scale {s}"
        np.testing.assert_allclose(q, ref, rtol=2e-4, atol=2e-4)
    bits = sum(float(q.sum(dtype=np.float64)) for q in mine)
    assert abs(bits - float(g["total_bits"])) <= 1e-5 * float(g["total_bits"])


@pytest.mark.parametrize("name", TRAINED_CASES)
def test_trained_checkpoint_cases_against_the_reference(name):
    """The same stage-by-stage comparison with weights trained by the reference's own training loop (sharper
    spreads than the hand-wired stand-ins): integer stages and header bit-exact, network outputs close, and on a
    host with the golden host's float behaviour every table digest and byte stream identical."""
    g = load_golden(name)
    cfg = oracle_config_for(name)
    sd = trained_state_dict()
    assert sum(v.size for v in sd.values()) == 196596                      # the reference's parameter count (exp_debug.log:101)
    d = O.StageDump()
    codec = O.OracleCodec(cfg, sd)
    bsl = codec.compress(g["rgb"], d)
    assert np.array_equal(d.ycocg, g["ycocg"]) and d.pad_int == int(g["pad_int"])
    for s in range(len(cfg.dwtlevels)):
        assert np.array_equal(d.planes[s], g[f"planes_{s}"])
    for j in range(9):
        assert bsl[0][j] == g[f"stream_0_{j}"].tobytes()
    for (s, b), p in d.params.items():
        np.testing.assert_allclose(p.reshape(60, -1)[:, ::7], g[f"params_{s}_{b}_sub"], rtol=1e-4, atol=1e-5)
    total = sum(len(x) for r in bsl for x in r)
    assert abs(total - int(g["total_bytes"])) <= 0.002 * int(g["total_bytes"]) + 2
    if host_matches_canary():
        for (s, b, c), t in d.tables.items():
            assert digest(t) == str(g[f"table_{s}_{b}_{c}_digest"])
        for i in range(1, len(bsl)):
            for j in range(9):
                assert bsl[i][j] == g[f"stream_{i}_{j}"].tobytes(), f"stream {i},{j}"
    assert np.array_equal(codec.decompress(bsl), g["rgb"])
