"""The training step on the GPU: `llicti_backward_dev` against the gradients of the unmodified reference
(golden fixtures) and against the oracle's autograd restatement on other batches; the model's autograd edge, an optimizer
step, and `mode: train` of the entry point."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, TRAIN_CASES, load_golden, oracle_config_for, train_state_dict
from oracle import llicti_oracle as O

pytestmark = pytest.mark.gpu

GRAD_TOL = 2e-4      # of a tensor's largest gradient: fp32 on both sides, other summation orders (atomics, tiles)


def _codec(L, ocfg, sd):
    from llicti_b200.codec import Codec, CodecConfig
    return Codec(CodecConfig(num_scales=len(ocfg.dwtlevels), chs=ocfg.chs, numerics=L.NUM_TORCH_CPU, cnn_impl=L.CNN_FP32), sd)


@pytest.fixture(scope="module")
def L():
    from llicti_b200 import _lib
    _lib.load()
    return _lib


def _compare(grads, ref, tol=GRAD_TOL):
    worst = {}
    for k, r in ref.items():
        q = grads[k].detach().cpu().numpy().reshape(r.shape)
        scale = float(np.abs(r).max())
        err = float(np.abs(q - r).max()) / (scale + 1e-30)
        worst[k] = err
    bad = {k: v for k, v in worst.items() if not v <= tol}
    assert not bad, f"gradients differ (relative to each tensor's largest entry): {bad}"
    return max(worst.values())


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_backward_equals_the_reference_gradients(L, name):
    g = load_golden(name)
    ocfg = oracle_config_for(name)
    sd = train_state_dict(name)
    codec = _codec(L, ocfg, sd)
    rgb = torch.from_numpy(g["rgb"]).cuda()
    sinfo = codec.forward_dev(rgb)
    loss = sum(float(s.double().sum()) for s in sinfo) / rgb.numel() * 3
    assert abs(loss - float(g["loss"])) <= 1e-4 * abs(loss), (loss, float(g["loss"]))
    gs = [torch.full_like(s, 3.0 / rgb.numel()) for s in sinfo]
    names = [k[5:] for k in g.files if k.startswith("grad/")]
    grads = codec.backward_dev(rgb, gs, names)                      # colour split and CNN recomputed
    _compare(grads, {k: g["grad/" + k] for k in names})
    sinfo2, kept = codec.train_forward_dev(rgb)                    # planes and network outputs kept from the forward pass
    assert all(torch.equal(a, b) for a, b in zip(sinfo, sinfo2))
    grads2 = codec.backward_dev(rgb, gs, names, kept=kept)
    _compare(grads2, {k: g["grad/" + k] for k in names})
    codec.close()


@pytest.mark.parametrize("cfgname,n,H,W", [("A", 3, 64, 96), ("B", 2, 60, 36), ("A", 1, 160, 160)], ids=["A_3x64x96", "B_2x60x36", "A_1x160x160"])
def test_backward_with_a_general_upstream_gradient(L, cfgname, n, H, W):
    """Positions that do not fill a 64-position tile, several tiles per CTA, signed upstream gradients (the LowerBound rule
    on the likelihood then blocks some of them) -- against the oracle's autograd."""
    ocfg = O.OracleConfig() if cfgname == "A" else O.OracleConfig(dwtlevels=(0, 1), chs=60)
    sd = O.jittered_state_dict(ocfg, seed=11)
    rgb_np = np.stack([O.synthetic_image(H, W, 70 + i, noise=3.0 * i) for i in range(n)])
    rng = np.random.default_rng(5)
    net = O.OracleNet(ocfg, sd)
    for t in net.sd.values():
        t.requires_grad_(True)
    x = torch.from_numpy(rgb_np.astype(np.float32) / np.float32(255.0))
    torch.set_num_threads(8)
    sinfos = O.self_informations_torch(ocfg, net, x)
    ups = [torch.from_numpy(rng.uniform(-0.5, 1.5, tuple(s.shape)).astype(np.float32)) for s in sinfos]
    sum((s * u).sum() for s, u in zip(sinfos, ups)).backward()
    ref = {k: t.grad.numpy() for k, t in net.sd.items()}
    codec = _codec(L, ocfg, sd)
    grads = codec.backward_dev(torch.from_numpy(rgb_np).cuda(), [u.cuda() for u in ups], list(ref))
    _compare(grads, ref)
    codec.close()


def test_training_step_through_the_model_interface(L):
    """The reference's step, verbatim (agents/llicti_agent.py:56-68): self_infos = model(x); loss; backward; clip; Adam.
    Gradients equal the oracle's; after the step the forward pass uses the updated weights (llicti_set_weights_dev)."""
    from llicti_b200 import LLICTI
    from llicti_b200.rate import TrainRLossList
    cfg = json.load(open(os.path.join(ROOT, "configs", "llicti_A.json")))
    ocfg = O.OracleConfig()
    sd = O.jittered_state_dict(ocfg, seed=1337)
    model = LLICTI(cfg, numerics=L.NUM_TORCH_CPU).to("cuda")
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model.train()
    opt = torch.optim.Adam([{"params": model.parameters(), "lr": 1e-4}])
    rgb_np = np.stack([O.synthetic_image(64, 64, 90 + i) for i in range(4)])
    x = (torch.from_numpy(rgb_np).float() / 255).cuda()
    losses = []
    for it in range(3):
        sinfos = model(x)
        assert all(s.requires_grad for s in sinfos)
        loss, table = TrainRLossList().forward(torch.numel(x), sinfos)
        loss.backward()
        if it == 0:
            o_loss, ref = O.train_loss_and_grads(ocfg, sd, rgb_np)
            assert abs(float(loss) - o_loss) <= 1e-4 * o_loss
            _compare({k: p.grad for k, p in model.named_parameters()}, ref)
        torch.nn.utils.clip_grad_value_(model.parameters(), clip_value=5.0)
        opt.step()
        opt.zero_grad()
        losses.append(float(loss))
    assert losses[2] < losses[0], losses
    # the forward pass of the training context sees the stepped weights: the oracle with the model's current state_dict
    sd_now = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    sinfos = model(x)
    want = O.forward_self_informations(ocfg, O.OracleNet(ocfg, sd_now), rgb_np[0])
    for s in range(5):
        np.testing.assert_allclose(sinfos[s][0].detach().cpu().numpy(), want[s], rtol=2e-3, atol=4e-3)
    # and the coding path (a fresh context: tcgen05 CNN) round-trips with them
    model.eval()
    bsl, _ = model.compress(x[:1])
    rec = model.decompres(bsl, "cuda")
    assert torch.equal(torch.round(rec * 255), torch.round(x[:1] * 255))


def test_set_weights_needs_the_fp32_context(L):
    from llicti_b200.codec import Codec, CodecConfig
    from llicti_b200._lib import LlictiError
    ocfg = O.OracleConfig()
    sd = O.synthetic_state_dict(ocfg)
    codec = Codec(CodecConfig(cnn_impl=L.CNN_TCGEN05), sd)
    t = {k: torch.from_numpy(v).cuda() for k, v in sd.items() if "conditional_prob_model" not in k}
    with pytest.raises(LlictiError, match="LLICTI_CNN_FP32"):
        codec.set_weights_dev(t)
    codec.close()


def test_main_train_mode(tmp_path):
    """`mode: train` through the entry point: two epochs over a few synthetic images, the 'tr' and 'va' tables, a
    checkpoint in the reference's format (optimizer, scheduler and logger states included) and a validation rate that
    went down."""
    from PIL import Image
    cfg = json.load(open(os.path.join(ROOT, "configs", "llicti_B.json")))
    tr, va = tmp_path / "train", tmp_path / "valid"
    tr.mkdir(), va.mkdir()
    for i in range(12):
        Image.fromarray(O.synthetic_image(72, 88, 200 + i).transpose(1, 2, 0)).save(tr / f"t{i}.png")
    for i in range(2):
        Image.fromarray(O.synthetic_image(64, 64, 300 + i).transpose(1, 2, 0)).save(va / f"v{i}.png")
    cfg.update({"mode": "train", "num_train_dirs": 1, "train_data_1": str(tr), "valid_data": str(va), "test_data": str(va),
                "patch_size": 64, "batch_size": 4, "patches_per_img": 1, "grad_acc_iters": 1, "loss_prnt_iters": 1000,
                "val_patch_size": 64, "val_batch_size": 2, "learning_rate": 1e-3, "max_epoch": 3, "validate_every": 1,
                "resume_training": False, "seed": 7})
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), str(cfg_path)], cwd=tmp_path, env=dict(os.environ, PYTHONPATH=ROOT),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    exp = os.path.join("experiments", cfg["multi_exp_name"], "exp_0")
    log = (tmp_path / exp / "logs" / "exp_debug.log").read_text()
    assert log.count("Train Epoch:") == 3 and log.count("Valid Epoch:") == 3, log[-3000:]
    import re
    totals = [float(v) for v in re.findall(r"\(\(([0-9.]+)\)\)", log)]
    valid = totals[1::2]
    assert len(valid) == 3 and valid[-1] < valid[0], totals
    ck = torch.load(tmp_path / exp / "checkpoints" / "checkpoint.pth.tar", map_location="cpu", weights_only=False)
    assert {"epoch", "iteration", "best_valid_loss", "state_dict", "optimizer", "scheduler", "train_logger", "valid_logger"} <= set(ck)
    assert ck["iteration"] == 9 and (tmp_path / exp / "checkpoints" / "model_best.pth.tar").exists()
    # resume (base.py:51-81 with resume_training): one more epoch from checkpoint.pth.tar, Adam's moments and the loggers restored
    cfg.update({"resume_training": True, "checkpoint_file": "checkpoint.pth.tar", "max_epoch": 4})
    cfg_path.write_text(json.dumps(cfg))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), str(cfg_path)], cwd=tmp_path, env=dict(os.environ, PYTHONPATH=ROOT),
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    ck2 = torch.load(tmp_path / exp / "checkpoints" / "checkpoint.pth.tar", map_location="cpu", weights_only=False)
    assert ck2["iteration"] == 12 and ck2["epoch"] == 4
    step = ck2["optimizer"]["state"][0]["step"]
    assert int(step) == 12, step                         # Adam continued from the restored state (9 steps) instead of restarting
    log = (tmp_path / exp / "logs" / "exp_debug.log").read_text()
    assert log.count("Train Epoch:") == 4 and "Checkpoint loaded successfully" in log


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_main_train_mode_two_ranks(tmp_path):
    """`mode: train` under torchrun: one process per GPU, each a share of every epoch, gradients averaged over NCCL; both
    ranks end with the same weights (rank 0's checkpoint, rank 1's dumped by the test hook)."""
    import socket
    from PIL import Image
    cfg = json.load(open(os.path.join(ROOT, "configs", "llicti_B.json")))
    tr, va = tmp_path / "train", tmp_path / "valid"
    tr.mkdir(), va.mkdir()
    for i in range(16):
        Image.fromarray(O.synthetic_image(72, 88, 500 + i).transpose(1, 2, 0)).save(tr / f"t{i}.png")
    for i in range(2):
        Image.fromarray(O.synthetic_image(64, 64, 600 + i).transpose(1, 2, 0)).save(va / f"v{i}.png")
    cfg.update({"mode": "train", "num_train_dirs": 1, "train_data_1": str(tr), "valid_data": str(va), "test_data": str(va),
                "patch_size": 64, "batch_size": 4, "patches_per_img": 1, "grad_acc_iters": 1, "loss_prnt_iters": 1000,
                "val_patch_size": 64, "val_batch_size": 2, "learning_rate": 1e-3, "max_epoch": 2, "validate_every": 1,
                "resume_training": False, "seed": 7})
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=ROOT, LLICTI_TEST_DUMP_WEIGHTS=str(tmp_path / "rank{rank}.pt"))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "main.py"), str(cfg_path)], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    exp = os.path.join("experiments", cfg["multi_exp_name"], "exp_0")
    ck = torch.load(tmp_path / exp / "checkpoints" / "checkpoint.pth.tar", map_location="cpu", weights_only=False)
    assert ck["iteration"] == 4                      # 16 images / 2 ranks / batch 4 = 2 steps per epoch, 2 epochs
    w0 = torch.load(tmp_path / "rank0.pt", map_location="cpu")
    w1 = torch.load(tmp_path / "rank1.pt", map_location="cpu")
    assert set(w0) == set(w1) == set(ck["state_dict"])
    for k in w0:
        assert torch.equal(w0[k], w1[k]), k          # the same averaged gradients, the same Adam: identical weights
        assert torch.equal(w0[k], ck["state_dict"][k]), k
