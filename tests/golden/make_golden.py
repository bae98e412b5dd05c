"""Generate golden vectors by running the UNMODIFIED reference files.

Run in the build container only (it needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

What it does
  1. puts /root/reference and oracle/refshims (stand-ins for compressai / torchac /
     easydict, see oracle/refshims/README.md) on sys.path and imports the reference's
     own `graphs.models.LLICTI_nets.LLICTI`;
  2. loads the deterministic synthetic weights (oracle.synthetic_state_dict -- the
     shipped checkpoint is absent) into the reference model;
  3. for each case runs the reference's colour transform, lazyDWT, get_params,
     get_cdfs, compress and decompres, checks the reference round-trips, and checks
     that oracle/llicti_oracle.py reproduces every stage bit-exactly on this host;
  4. writes tests/golden/<case>.npz: the small integer artefacts verbatim (image,
     planes, header, all byte streams), strided subsamples of the big float/int
     tables, and sha256 digests of the full arrays.

The coder under the reference's `torchac.*` calls is oracle/torchac_port.c (real
torchac is not installable here): byte streams are "reference model code + restated
coder" -- see the parity note in that file.
"""
import hashlib
import json
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "refshims"))
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import llicti_oracle as O  # noqa: E402
from easydict import EasyDict  # noqa: E402  (shim)
from graphs.models.LLICTI_nets import LLICTI  # noqa: E402  (the reference)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_config(name):
    with open(os.path.join(REF, "configs", name)) as f:
        return EasyDict(json.load(f))


def make_image(kind, H, W, idx):
    if kind == "photo":
        return O.synthetic_image(H, W, idx)
    if kind == "photo_lownoise":
        return O.synthetic_image(H, W, idx, noise=0.7)
    if kind == "const":
        img = np.empty((3, H, W), dtype=np.uint8)
        img[0], img[1], img[2] = 200, 31, 97
        return img
    if kind == "noise":
        return np.random.default_rng(idx).integers(0, 256, size=(3, H, W), dtype=np.uint8)
    if kind == "checker":
        yy, xx = np.mgrid[0:H, 0:W]
        v = (((yy + xx) & 1) * 255).astype(np.uint8)
        return np.stack([v, 255 - v, v])
    raise ValueError(kind)


CASES = [
    # name, config, kind, H, W
    ("a_photo_33x47", "llicti_A.json", "photo", 33, 47),
    ("a_photo_53x77", "llicti_A.json", "photo", 53, 77),
    ("a_photo_64x96", "llicti_A.json", "photo", 64, 96),
    ("a_const_40x72", "llicti_A.json", "const", 40, 72),
    ("a_noise_32x64", "llicti_A.json", "noise", 32, 64),
    ("a_checker_35x32", "llicti_A.json", "checker", 35, 32),
    ("b_photo_64x96", "llicti_B.json", "photo", 64, 96),
    ("b_photo_37x53", "llicti_B.json", "photo", 37, 53),
]

# the same pipeline with weights TRAINED by the reference's own `mode: train` (tools/train_reference_ckpt.py ->
# tests/golden/ckpt_A_trained.npz): spreads down to the 0.11-level clamp where the content allows, i.e. the regime in
# which rounding of the predicted means moves the rate (the hand-wired stand-in weights keep spreads at 1.5-10 levels)
TRAINED_CASES = [
    ("t_photo_64x96", "llicti_A.json", "photo_lownoise", 64, 96),
    ("t_photo_53x77", "llicti_A.json", "photo", 53, 77),
]


def trained_state_dict():
    with np.load(os.path.join(HERE, "ckpt_A_trained.npz")) as z:
        return {k: z[k] for k in z.files}


SUB = 7    # stride of the position subsample stored for the float parameter arrays
TSUB = 37  # stride of the row subsample stored for the integer CDF tables


def run_case(name, cfg_name, kind, H, W, idx, sd=None):
    cfg = ref_config(cfg_name)
    ocfg = O.OracleConfig.from_dict(cfg)
    if sd is None:
        sd = O.synthetic_state_dict(ocfg, seed=1337)
    torch.manual_seed(0)
    model = LLICTI(cfg).eval()
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected and all("conditional_prob_model" in k for k in missing), (missing, unexpected)

    rgb = make_image(kind, H, W, idx)
    x = torch.from_numpy(rgb.astype(np.float32) / np.float32(255.0))[None]  # what ToTensor() yields
    out = {"rgb": rgb, "config": np.array(cfg_name)}
    S, M = len(cfg.dwtlevels), cfg.num_mixtures

    with torch.no_grad():
        # --- stage dumps through the reference's own functions -------------------
        ycc = model.get_YCoCg_R_from_RGB__intOps(x.clone())
        minmax = [0, ycc[:, 1].min().item(), ycc[:, 2].min().item(), 255, ycc[:, 1].max().item(), ycc[:, 2].max().item()]
        xc = ycc.clone()
        xc[:, 0] = xc[:, 0] - 127
        xf = xc / 255
        y_list, flags, pad_int = model.lazyDWT(xf, levels=model.list_scales, clrchs=3, clrjnt=2, pad=True)
        out["ycocg"] = ycc[0].numpy()
        out["minmax"] = np.array(minmax, dtype=np.int16)
        out["pad_int"] = np.array(pad_int)
        out["pad_flags"] = np.array(flags, dtype=np.uint8)
        for s in range(S):
            out[f"planes_{s}"] = torch.round(y_list[s][0] * 255).to(torch.int16).numpy()

        # oracle, same stages
        dump = O.StageDump()
        codec = O.OracleCodec(ocfg, sd)
        o_bsl = codec.compress(rgb, dump)
        assert np.array_equal(dump.ycocg, out["ycocg"]), "ycocg"
        assert dump.minmax == minmax and dump.pad_int == pad_int and dump.pad_flags == [list(map(bool, f)) for f in flags]
        for s in range(S):
            assert np.array_equal(dump.planes[s], out[f"planes_{s}"]), f"planes {s}"

        shifts = [127, -minmax[1], -minmax[2]]
        for s in range(S - 1, -1, -1):
            yl = y_list[s]
            for b in range(3):
                mdl = model.entropymodel.entmdls_scale_band[0][b]
                params = mdl.get_params(yl[:, 0:3 * (b + 1)])
                p_np = params[0].numpy().copy()
                assert np.array_equal(p_np, dump.params[(s, b)]), f"params {s},{b}"
                out[f"params_{s}_{b}_digest"] = np.array(digest(p_np))
                out[f"params_{s}_{b}_sub"] = p_np.reshape(12 * M, -1)[:, ::SUB].copy()
                aw, bw, dw = params[:, 9 * M:10 * M], params[:, 10 * M:11 * M], params[:, 11 * M:12 * M]
                padH, padW = flags[s]
                for clr in range(3):
                    sig = params[:, clr * M:(clr + 1) * M]
                    mu = params[:, (3 + clr) * M:(4 + clr) * M]
                    wt = params[:, (6 + clr) * M:(7 + clr) * M]
                    if clr == 1:
                        mu += aw * yl[:, 3 * (b + 1):3 * (b + 1) + 1]
                    elif clr == 2:
                        mu += bw * yl[:, 3 * (b + 1):3 * (b + 1) + 1] + dw * yl[:, 3 * (b + 1) + 1:3 * (b + 1) + 2]
                    lo = -127 if clr == 0 else minmax[clr]
                    hi = 128 if clr == 0 else minmax[3 + clr]
                    tab = mdl.get_cdfs(sig, mu, wt, clrch=clr, int_cdf=True, minVal=lo, maxVal=hi)[0, 0].numpy()
                    ch, cw = O.crop_shape(b, yl.shape[2], yl.shape[3], padH, padW)
                    t = np.ascontiguousarray(tab[:ch, :cw]).reshape(ch * cw, -1)
                    assert np.array_equal(t, dump.tables[(s, b, clr)]), f"table {s},{b},{clr}"
                    out[f"table_{s}_{b}_{clr}_digest"] = np.array(digest(t))
                    out[f"table_{s}_{b}_{clr}_sub"] = t[::TSUB].copy()

        # --- the reference's own compress / decompres ----------------------------------
        bsl, _ = model.compress(x.clone())
        rec = model.decompres(bsl, torch.device("cpu"))
        err = ((x - rec) * 255).abs().max().item()
        assert err < 0.5, f"reference did not round-trip: {err}"
        assert len(bsl) == 1 + S and all(len(r) == 9 for r in bsl)
        for i, row in enumerate(bsl):
            for j, blob in enumerate(row):
                out[f"stream_{i}_{j}"] = np.frombuffer(blob, dtype=np.uint8).copy()
                assert blob == o_bsl[i][j], f"stream {i},{j} differs between reference and oracle"
        assert np.array_equal(codec.decompress(bsl), rgb)
        total = sum(len(b_) for r in bsl for b_ in r)
        out["total_bytes"] = np.array(total)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {H}x{W} {cfg_name} bytes={total} bpsp={total * 8 / rgb.size:.3f} pad_int={pad_int} "
          f"minmax={minmax} -> OK (reference == oracle, lossless)")


FORWARD_CASES = [
    # name, config, kind, H, W, image index  (H and W multiples of 2^S: the un-padded lazyDWT of forward())
    ("fwd_a_photo_64x96", "llicti_A.json", "photo", 64, 96, 2),
    ("fwd_a_noise_32x64", "llicti_A.json", "noise", 32, 64, 4),
    ("fwd_a_checker_32x32", "llicti_A.json", "checker", 32, 32, 5),
    ("fwd_b_photo_64x96", "llicti_B.json", "photo", 64, 96, 6),
    ("fwd_b_photo_36x52", "llicti_B.json", "photo", 36, 52, 7),
]


def run_forward_case(name, cfg_name, kind, H, W, idx):
    """LLICTI.forward (the rate-estimation path of validate / training) of the unmodified reference:
    self-informations per scale, stored whole (they are small), and checked bit-exactly against the oracle's
    restatement (oracle.forward_self_informations)."""
    cfg = ref_config(cfg_name)
    ocfg = O.OracleConfig.from_dict(cfg)
    sd = O.synthetic_state_dict(ocfg, seed=1337)
    torch.manual_seed(0)
    model = LLICTI(cfg).eval()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    rgb = make_image(kind, H, W, idx)
    x = torch.from_numpy(rgb.astype(np.float32) / np.float32(255.0))[None]
    with torch.no_grad():
        ref = model.forward(x.clone())
    mine = O.forward_self_informations(ocfg, O.OracleNet(ocfg, sd), rgb)
    out = {"rgb": rgb, "config": np.array(cfg_name)}
    for s, (r, q) in enumerate(zip(ref, mine)):
        assert np.array_equal(r[0].numpy(), q), f"{name}: self-informations of scale {s} differ between reference and oracle"
        out[f"sinfo_{s}"] = q
    bits = sum(float(q.sum(dtype=np.float64)) for q in mine)
    out["total_bits"] = np.array(bits)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {H}x{W} {cfg_name} estimated {bits / (H * W):.3f} bpp -> OK (reference forward == oracle)")


TRAIN_CASES = [
    # name, config, B, H, W, first image index: one backward pass of the reference's training step
    ("train_a_2x32x32", "llicti_A.json", 2, 32, 32, 30),
    ("train_b_3x24x40", "llicti_B.json", 3, 24, 40, 33),
    ("train_t_2x64x32", "llicti_A.json", 2, 64, 32, 36),      # weights trained by the reference (ckpt_A_trained.npz): spreads at the clamp
]


def run_train_case(name, cfg_name, B, H, W, idx):
    """The backward pass of the unmodified reference's training step (agents/llicti_agent.py:52-61): model.train(),
    self_infos = model(x), TrainRLossList, loss.backward().  Stored: the batch, the loss and every parameter's gradient;
    checked against the oracle's restatement (oracle.train_loss_and_grads)."""
    from graphs.losses.rate_dist import TrainRLossList
    cfg = ref_config(cfg_name)
    ocfg = O.OracleConfig.from_dict(cfg)
    # generic weights: no pre-activation sits exactly on a ReLU kink
    sd = trained_state_dict() if name.startswith("train_t_") else O.jittered_state_dict(ocfg, seed=1337)
    torch.manual_seed(0)
    model = LLICTI(cfg).train()
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    rgb = np.stack([make_image("photo" if i % 2 == 0 else "noise", H, W, idx + i) for i in range(B)])
    x = torch.from_numpy(rgb.astype(np.float32) / np.float32(255.0))
    sinfos = model(x.clone())
    loss, rate1_list = TrainRLossList().forward(torch.numel(x), sinfos)
    loss.backward()
    ref = {k: p.grad.numpy() for k, p in model.named_parameters() if p.grad is not None}
    o_loss, mine = O.train_loss_and_grads(ocfg, sd, rgb)
    mine = {k: v for k, v in mine.items() if k in ref}
    assert set(ref) == set(mine), (set(ref) ^ set(mine))
    worst = 0.0
    for k in ref:
        err = float(np.abs(ref[k] - mine[k]).max()) / (float(np.abs(ref[k]).max()) + 1e-30)
        worst = max(worst, err)
        assert err < 1e-5, f"{name}: gradient of {k} differs between reference and oracle ({err:.3e})"
    assert abs(o_loss - float(loss.item())) < 1e-6 * abs(o_loss)
    out = {"rgb": rgb, "config": np.array(cfg_name), "loss": np.array(float(loss.item()))}
    for k, v in ref.items():
        out["grad/" + k] = v.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {B}x{H}x{W} {cfg_name} loss {float(loss.item()):.4f} bpp, {len(ref)} gradients, "
          f"worst relative difference reference vs oracle {worst:.2e} -> OK")


def canary():
    """Bit patterns of the two host-dependent float primitives (vector erfc, reduction
    order); tests skip the bit-exact float checks when the running host disagrees."""
    x = torch.linspace(-6, 6, 4001, dtype=torch.float32)
    e = torch.erfc(x)
    w = torch.from_numpy(np.random.default_rng(5).random((1, 37, 53, 1, 5), dtype=np.float32))
    wp = w.permute(0, 4, 1, 2, 3).contiguous().permute(0, 2, 3, 4, 1)   # layout of the reference's weights
    s = torch.sum(wp, dim=4)
    np.savez_compressed(os.path.join(HERE, "canary.npz"), erfc_digest=np.array(digest(e.numpy())),
                        sum_digest=np.array(digest(s.numpy())))


if __name__ == "__main__":
    torch.set_num_threads(8)
    canary()
    if "--only-train-step" in sys.argv:
        for c in TRAIN_CASES:
            run_train_case(*c)
        sys.exit(0)
    if "--only-forward" not in sys.argv and "--only-trained" not in sys.argv:
        for i, c in enumerate(CASES):
            run_case(*c, idx=i)
    if "--only-forward" not in sys.argv:
        for i, c in enumerate(TRAINED_CASES):
            run_case(*c, idx=20 + i, sd=trained_state_dict())
    if "--only-trained" in sys.argv:
        sys.exit(0)
    for c in FORWARD_CASES:
        run_forward_case(*c)
    for c in TRAIN_CASES:
        run_train_case(*c)
    sz = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print("fixtures total bytes:", sz)
