"""CPU-side checks of the C ABI and the host logic (no compute calls: no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import llicti_oracle as O


def test_library_exports_every_declared_symbol(built_lib):
    from llicti_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "llicti.h")).read()
    declared = set(re.findall(r"LLICTI_API[^;(]*?\b(llicti_\w+)\s*\(", hdr))
    assert len(declared) >= 24 and {"llicti_status", "llicti_forward_dev", "llicti_cnn_operands"} <= declared
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for name in declared:
        assert hasattr(built_lib, name), name
    assert built_lib.llicti_abi_version() == 1


def test_struct_layouts_match_header(built_lib):
    from llicti_b200 import _lib as L
    assert C.sizeof(L.Config) == 32
    assert C.sizeof(L.Weights) == 8 * 24
    # int32 x (3 + 4*8 + 1 + 3*24) = 108 -> 432 bytes, then 4 x int64
    assert C.sizeof(L.Geom) == 432 + 32


@pytest.mark.parametrize("H,W,S", [(512, 768, 5), (1356, 2040, 5), (2160, 3840, 5), (512, 512, 5), (33, 47, 5),
                                   (53, 77, 5), (35, 32, 5), (17, 17, 5), (64, 96, 2), (37, 53, 2), (255, 257, 3)])
def test_geometry_matches_oracle_pyramid(built_lib, H, W, S):
    from llicti_b200 import _lib as L
    cfg = L.Config(S, 88, 5, 0, 0, 0, 0, 0)
    g = L.Geom()
    assert built_lib.llicti_geometry(C.byref(cfg), H, W, C.byref(g)) == 0
    planes, flags, pad_int = O.pyramid_split(np.zeros((3, H, W), dtype=np.int16), tuple(range(S)))
    assert g.pad_int == pad_int
    sym = 0
    for s in range(S):
        assert (g.Hs[s], g.Ws[s]) == planes[s].shape[1:]
        assert [bool(g.padH[s]), bool(g.padW[s])] == flags[s]
        for b in range(3):
            ch, cw = O.crop_shape(b, g.Hs[s], g.Ws[s], *flags[s])
            assert (g.crop_h[s][b], g.crop_w[s][b]) == (ch, cw)
            sym += 3 * ch * cw
    assert g.symbols == sym
    assert g.positions == sum(p.shape[1] * p.shape[2] for p in planes)


def test_survey_geometry_table(built_lib):
    from llicti_b200 import _lib as L
    cfg = L.Config(5, 88, 5, 0, 0, 0, 0, 0)
    g = L.Geom()
    for (H, W), (pos, sym, pad) in {(512, 768): (130944, 1178496, 0), (1356, 2040): (921432, 8290464, 38),
                                    (2160, 3840): (2762160, 24858720, 2), (512, 512): (87296, 785664, 0)}.items():
        assert built_lib.llicti_geometry(C.byref(cfg), H, W, C.byref(g)) == 0
        assert (g.positions, g.symbols, g.pad_int) == (pos, sym, pad)


@pytest.mark.parametrize("sub_len", [64, 2048])
def test_substream_counts_match_oracle(built_lib, sub_len):
    from llicti_b200 import _lib as L
    cfg = L.Config(5, 88, 5, sub_len, 0, 0, 0, 0)
    g = L.Geom()
    assert built_lib.llicti_geometry(C.byref(cfg), 512, 768, C.byref(g)) == 0
    tot = 0
    for s in range(5):
        for b in range(3):
            n = g.crop_h[s][b] * g.crop_w[s][b]
            assert g.num_sub[s][b] == O.num_substreams(n, O.scale_sub_len(sub_len, s))
            tot += 3 * g.num_sub[s][b]
    assert g.substreams == tot


def test_error_behaviour_without_device(built_lib):
    from llicti_b200 import _lib as L
    g = L.Geom()
    bad = L.Config(5, 88, 5, 0, 0, 0, 0, 0)
    assert built_lib.llicti_geometry(C.byref(bad), 8, 8, C.byref(g)) == L.E_ARG          # too small for 5 scales
    assert b"too small" in built_lib.llicti_last_error()
    assert built_lib.llicti_geometry(C.byref(bad), 512, 8161 * 2, C.byref(g)) == L.E_ARG  # coarsest > 255 (uint8 header)
    bad = L.Config(5, 88, 4, 0, 0, 0, 0, 0)
    assert built_lib.llicti_geometry(C.byref(bad), 64, 64, C.byref(g)) == L.E_ARG
    if built_lib.llicti_device_count() == 0:
        ctx = C.c_void_p()
        w = L.Weights()
        ok = L.Config(5, 88, 5, 0, 0, 0, 0, 0)
        assert built_lib.llicti_create(C.byref(ok), C.byref(w), C.byref(ctx)) == L.E_NODEVICE
        assert b"no CPU fallback" in built_lib.llicti_last_error()


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from llicti_b200 import Codec, CodecConfig
    cfg = O.OracleConfig()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Codec(CodecConfig(), O.synthetic_state_dict(cfg))


def test_container_assemble_parse_round_trip():
    from llicti_b200 import container
    rng = np.random.default_rng(0)
    S, n, H, W = 2, 3, 37, 53
    planes, flags, pad_int = O.pyramid_split(np.zeros((3, H, W), dtype=np.int16), (0, 1))
    h_last, w_last = planes[-1].shape[1:]
    rgb = rng.integers(0, 256, size=(n, 3, H, W), dtype=np.uint8)
    lens = rng.integers(0, 50, size=n * 9 * S)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    blob = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    mm = rng.integers(-255, 256, size=(n, 6)).astype(np.int16)
    for sub_len in (0, 512):
        bsls = container.assemble(S, sub_len, h_last, w_last, pad_int, rgb, blob, off, mm)
        assert len(bsls) == n and all(len(b) == S + 1 and all(len(r) == 9 for r in b) for b in bsls)
        assert bsls[0][0][0] == bytes([S, h_last, w_last])
        assert bsls[0][0][3] == np.ascontiguousarray(rgb[0, :, ::4, ::4]).tobytes()
        blob2, off2, mm2, x00, n2, H2, W2 = container.parse(S, sub_len, bsls)
        assert (n2, H2, W2) == (n, H, W)
        assert np.array_equal(blob2, blob) and np.array_equal(off2, off) and np.array_equal(mm2, mm)
        assert np.array_equal(x00, rgb[:, :, ::4, ::4])
        with pytest.raises(ValueError):
            container.parse(S, sub_len + 1, bsls)
    with pytest.raises(ValueError):
        container.parse(3, 0, container.assemble(S, 0, h_last, w_last, pad_int, rgb, blob, off, mm))


def test_header_matches_oracle_header():
    from llicti_b200 import container
    cfg = O.OracleConfig()
    img = O.synthetic_image(53, 77, 2)
    o = O.OracleCodec(cfg, O.synthetic_state_dict(cfg), sub_len=128)
    d = O.StageDump()
    bsl = o.compress(img, d)
    mm = np.array([d.minmax], dtype=np.int16)
    h_last, w_last = d.planes[-1].shape[1:]
    off = np.zeros(46, dtype=np.uint64)
    mine = container.assemble(5, 128, h_last, w_last, d.pad_int, img[None], np.zeros(1, np.uint8), off, mm)[0]
    assert mine[0] == bsl[0]
    assert container.image_size_from_header(5, h_last, w_last, d.pad_int) == (53, 77)


def test_config_validation():
    import json
    from llicti_b200 import CodecConfig
    for name, chs, S in (("llicti_A.json", 88, 5), ("llicti_B.json", 60, 2)):
        cfg = json.load(open(os.path.join(ROOT, "configs", name)))
        cc = CodecConfig.from_json_dict(cfg)
        assert (cc.chs, cc.num_scales) == (chs, S)
        assert cfg["mode"] == "eval_model" and cfg["agent"] == "LLICTIAgent"
        bad = dict(cfg, clr_joint_mode=0)
        with pytest.raises(ValueError):
            CodecConfig.from_json_dict(bad)


def test_model_state_dict_names_match_reference_layout():
    import json
    import torch
    from llicti_b200 import LLICTI
    cfg = json.load(open(os.path.join(ROOT, "configs", "llicti_A.json")))
    m = LLICTI(cfg)
    names = dict(m.named_parameters())
    assert len(names) == 24
    assert sum(p.numel() for p in names.values()) == 196596
    ocfg = O.OracleConfig()
    sd = {k: torch.from_numpy(v) for k, v in O.synthetic_state_dict(ocfg).items()}
    assert set(sd) == set(names)
    # compressai's extra buffers in a reference checkpoint must not break loading
    sd["entropymodel.entmdls_scale_band.0.0.conditional_prob_model._offset"] = torch.zeros(0, dtype=torch.int32)
    m.load_state_dict(sd)
    assert torch.equal(m.state_dict()["entropymodel.entmdls_scale_band.0.1.layer0_11_01.weight"],
                       sd["entropymodel.entmdls_scale_band.0.1.layer0_11_01.weight"])


def test_llicti_file_round_trip_and_validation():
    """.llicti container: the reference's bytestream_list, length-prefixed."""
    from llicti_b200 import container, fileformat
    rng = np.random.default_rng(5)
    S, H, W = 2, 37, 53
    planes, flags, pad_int = O.pyramid_split(np.zeros((3, H, W), dtype=np.int16), (0, 1))
    h_last, w_last = planes[-1].shape[1:]
    rgb = rng.integers(0, 256, size=(1, 3, H, W), dtype=np.uint8)
    lens = rng.integers(0, 40, size=9 * S)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    blob = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    mm = rng.integers(-255, 256, size=(1, 6)).astype(np.int16)
    for sub_len in (0, 256):
        bsl = container.assemble(S, sub_len, h_last, w_last, pad_int, rgb, blob, off, mm)[0]
        data = fileformat.dumps(bsl, sub_len, H, W)
        bsl2, sub2, H2, W2 = fileformat.loads(data)
        assert (sub2, H2, W2) == (sub_len, H, W) and bsl2 == [list(map(bytes, r)) for r in bsl]
        assert len(data) == 21 + 4 * 9 * (S + 1) + sum(len(e) for r in bsl for e in r)
        for bad in (data[:-1], data + b"x", b"XLICTI" + data[6:], data[:6] + bytes([9]) + data[7:]):
            with pytest.raises(ValueError):
                fileformat.loads(bad)
        with pytest.raises(ValueError):
            fileformat.dumps(bsl, sub_len + 1, H, W)
        with pytest.raises(ValueError):
            fileformat.loads(fileformat.dumps(bsl, sub_len, H + 1, W))


def test_synthetic_inputs_of_bench_equal_the_oracle_generators():
    """bench.py and the tools take their synthetic images / weights from llicti_b200.synth (the product path imports
    nothing from oracle/); both generators must produce the same bytes."""
    from llicti_b200 import synth
    for cfg in (O.OracleConfig(), O.OracleConfig(dwtlevels=(0, 1), chs=60)):
        a = O.synthetic_state_dict(cfg)
        b = synth.synthetic_state_dict(cfg.chs, cfg.num_mixtures, cfg.evens, cfg.odds)
        assert a.keys() == b.keys()
        for k in a:
            assert a[k].dtype == b[k].dtype and np.array_equal(a[k], b[k]), k
    for H, W, i in ((33, 47, 0), (64, 96, 5), (17, 17, 9)):
        assert np.array_equal(O.synthetic_image(H, W, i), synth.synthetic_image(H, W, i))


def test_product_sources_do_not_import_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may use oracle/."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "llicti_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "oracle/" in text and f.endswith((".cu", ".cuh", ".h")) and "#include" in text and re.search(r'#include\s+"[^"]*oracle', text):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    bench = open(os.path.join(root, "bench.py")).read()
    body = bench[bench.index("# B200 arm"):bench.index("def main(")]
    assert not re.search(r"^\s*(from|import)\s+oracle\b", body, re.M), "bench.py's product arm imports the oracle"
    assert "cpu_oracle_pass(" not in body, "bench.py's product arm runs the CPU oracle in-process (it belongs in the reference arm's process)"
    main = open(os.path.join(root, "main.py")).read()
    assert not re.search(r"^\s*(from|import)\s+oracle\b", main, re.M)


def test_container_fingerprint_and_checksum():
    """A stream records what its decoder must share with the encoder (CNN implementation, numerics profile, weights)
    and the checksum of its image; a codec with another fingerprint refuses it.  Streams without a fingerprint
    (the reference's own header row leaves slots 4-8 empty) are accepted."""
    from llicti_b200 import container
    rng = np.random.default_rng(3)
    S, H, W = 2, 20, 24
    rgb = rng.integers(0, 256, size=(1, 3, H, W), dtype=np.uint8)
    off = np.arange(9 * S + 1, dtype=np.uint64) * 3
    blob = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    mm = np.array([[0, -5, -7, 255, 9, 11]], dtype=np.int16)
    fp = container.fingerprint(1, 0, 0xDEADBEEF)
    bsl = container.assemble(S, 0, 5, 6, 0, rgb, blob, off, mm, fp=fp, checksum=True)[0]
    assert bsl[0][5] == fp and len(bsl[0][6]) == 4 and bsl[0][7:] == [b"", b""]
    import zlib
    assert container.image_checksum(bsl[0]) == zlib.crc32(rgb[0].tobytes())
    container.parse(S, 0, [bsl], fp=fp)                                       # same codec: fine
    with pytest.raises(ValueError, match="cnn=tcgen05/bf16.*cnn=fp32"):
        container.parse(S, 0, [bsl], fp=container.fingerprint(0, 0, 0xDEADBEEF))
    with pytest.raises(ValueError, match="weights crc32"):
        container.parse(S, 0, [bsl], fp=container.fingerprint(1, 0, 0x12345678))
    plain = container.assemble(S, 0, 5, 6, 0, rgb, blob, off, mm)[0]          # the reference's header row: nothing to check
    assert plain[0][4:] == [b""] * 5 and container.image_checksum(plain[0]) is None
    container.parse(S, 0, [plain], fp=fp)
    bad = [list(r) for r in bsl]
    bad[0][5] = b"\x07abc"
    with pytest.raises(ValueError):
        container.parse(S, 0, [bad], fp=fp)


def test_rate_table_text_is_the_references():
    """RateLogger.display prints the reference's table (loggers/rate.py:120-168): the expected strings below were
    produced by the unmodified reference's text_log_list on the same numbers (tools: see the test body)."""
    from datetime import datetime
    from llicti_b200.rate import RateLogger
    rate = (np.arange(3 * 9, dtype=np.float64).reshape(3, 9) + 1) / 8
    now = datetime(2020, 1, 2, 3, 4, 5)
    te = RateLogger.format_table(7, rate, 0.0, "te", now)
    assert te == ("   Test Epoch:   7  Rates: hdr -> 0.12+0.25+0.38(b0=0.750) 0.50+0.62+0.75(b1=1.875) 0.88+1.00+1.12(b2=3.000) (hd=5.625) \n"
                  "                                   scl0-> 1.25+1.38+1.50(b0=4.125) 1.62+1.75+1.88(b1=5.250) 2.00+2.12+2.25(b2=6.375) (s0=15.750) \n"
                  "                                   scl1-> 2.38+2.50+2.62(b0=7.500) 2.75+2.88+3.00(b1=8.625) 3.12+3.25+3.38(b2=9.750) (s1=25.875) "
                  "((47.250))  (03:04:05)")
    va = RateLogger.format_table(7, rate[:1], 0.0, "va", now)
    assert va == "  Valid Epoch:   7  Rates: scl0-> 0.12+0.25+0.38(b0=0.750) 0.50+0.62+0.75(b1=1.875) 0.88+1.00+1.12(b2=3.000) (s0=5.625) ((5.625))  (03:04:05)"
    if os.path.exists("/root/reference/loggers/rate.py"):                      # build container: against the reference itself
        import importlib.util
        import types
        spec = importlib.util.spec_from_file_location("ref_rate", "/root/reference/loggers/rate.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ref = mod.RateLogger()
        lines = []
        ref.logger = types.SimpleNamespace(info=lines.append)
        ref._get_time_now_str = lambda: now.strftime("%H:%M:%S")
        for typ in ("te", "tr", "va", "it"):
            lines.clear()
            ref.text_log_list(7, rate, 0.001, typ)
            assert lines[0] == RateLogger.format_table(7, rate, 0.001, typ, now), typ
