"""`python main.py <config.json>` in eval_model mode on a scratch experiment directory."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import llicti_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg_name,ocfg", [("llicti_A.json", O.OracleConfig()),
                                           ("llicti_B.json", O.OracleConfig(dwtlevels=(0, 1), chs=60))])
def test_main_eval_model(tmp_path, cfg_name, ocfg):
    from PIL import Image
    cfg = json.load(open(os.path.join(ROOT, "configs", cfg_name)))
    data = tmp_path / "data"
    data.mkdir()
    for i, (h, w) in enumerate([(64, 96), (53, 77)]):
        Image.fromarray(O.synthetic_image(h, w, i).transpose(1, 2, 0)).save(data / f"img{i}.png")
    cfg["test_data"] = str(data)
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    exp = os.path.join("experiments", cfg["multi_exp_name"], "exp_0")
    ckdir = tmp_path / exp / "checkpoints"
    ckdir.mkdir(parents=True)
    sd = {k: torch.from_numpy(v) for k, v in O.synthetic_state_dict(ocfg).items()}
    torch.save({"epoch": 1, "iteration": 2, "best_valid_loss": np.float64(1.0), "state_dict": sd},
               ckdir / "model_best.pth.tar")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), str(cfg_path)], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    log = (tmp_path / exp / "logs" / "exp_debug.log").read_text()
    assert log.count("(Check: Decoded img matches original)") == 2, log[-2000:]
    assert "Checkpoint loaded successfully" in log
    # the shipped configs run the product CNN (tcgen05), and its kernel class really launched
    import re
    assert "cnn_impl=1 (tcgen05" in log, log[-2000:]
    m = re.search(r"cnn\[tcgen05\]=(\d+)", log)
    assert m and int(m.group(1)) >= 2 * 2 * 3 * len(ocfg.dwtlevels), log[-2000:]
    assert "Test Epoch:" in log and "Rates: hdr ->" in log and "(hd=" in log       # the reference's rate table text
    assert (ckdir / "checkpoint.pth.tar").exists()


def test_main_validate_mode(tmp_path):
    """`mode: validate` (the reference's validate(): forward() rate estimate over centre crops of the validation images) runs
    through the entry point and logs the 'va' table; its total is the oracle's rate estimate of the same crops."""
    from PIL import Image
    ocfg = O.OracleConfig()
    cfg = json.load(open(os.path.join(ROOT, "configs", "llicti_A.json")))
    data = tmp_path / "valid"
    data.mkdir()
    imgs = [O.synthetic_image(80, 112, 20 + i) for i in range(3)]
    for i, im in enumerate(imgs):
        Image.fromarray(im.transpose(1, 2, 0)).save(data / f"v{i}.png")
    cfg.update({"mode": "validate", "valid_data": str(data), "val_patch_size": 64, "val_batch_size": 2, "test_data": str(data)})
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    exp = os.path.join("experiments", cfg["multi_exp_name"], "exp_0")
    ckdir = tmp_path / exp / "checkpoints"
    ckdir.mkdir(parents=True)
    sd_np = O.synthetic_state_dict(ocfg)
    torch.save({"epoch": 1, "iteration": 2, "best_valid_loss": np.float64(1.0),
                "state_dict": {k: torch.from_numpy(v) for k, v in sd_np.items()}}, ckdir / "model_best.pth.tar")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), str(cfg_path)], cwd=tmp_path, env=dict(os.environ, PYTHONPATH=ROOT),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    log = (tmp_path / exp / "logs" / "exp_debug.log").read_text()
    assert "Valid Epoch:" in log and "Rates: scl0->" in log, log[-2000:]
    import re
    total = float(re.findall(r"\(\(([0-9.]+)\)\)", log)[-1])
    # the oracle's forward() on the same centre crops (64 x 64: a multiple of the 32 the agent pads to)
    want, net = [], O.OracleNet(ocfg, sd_np)
    for im in imgs:
        crop = np.ascontiguousarray(im[:, 8:72, 24:88])
        sinfo = O.forward_self_informations(ocfg, net, crop)
        want.append(sum(float(s.sum()) for s in sinfo) / crop.size * 3)
    # batches of (2, 1) images: the table averages the per-batch rates
    per_batch = [(want[0] + want[1]) / 2, want[2]]
    assert abs(total - sum(per_batch) / 2) < 0.02 * total, (total, want)


@pytest.mark.gpu
@pytest.mark.parametrize("sub_len", [0, 512])
def test_cli_encode_decode_files(tmp_path, sub_len):
    """PNG -> .llicti -> PNG through the command-line front end is lossless."""
    from PIL import Image
    from llicti_b200 import cli
    img = O.synthetic_image(75, 109, 4)
    src, mid, dst = tmp_path / "in.png", tmp_path / "x.llicti", tmp_path / "out.png"
    Image.fromarray(np.ascontiguousarray(img.transpose(1, 2, 0)), "RGB").save(src)
    cfg = os.path.join(ROOT, "configs", "llicti_A.json")
    assert cli.main(["encode", str(src), str(mid), "--config", cfg, "--sub-len", str(sub_len)]) == 0
    assert cli.main(["decode", str(mid), str(dst), "--config", cfg]) == 0
    assert np.array_equal(np.asarray(Image.open(dst)).transpose(2, 0, 1), img)


@pytest.mark.parametrize("mode,needle", [("flops_est", "Computational complexity:"), ("model_size", "model param+buffer=total size"),
                                         ("test", "Checkpoint loaded successfully")])
def test_main_auxiliary_modes(tmp_path, mode, needle):
    """The reference's remaining agent modes run through the entry point (flops_est: closed-form MACs of one 512 x 512
    forward, cfg A: 64 + 16 + 4 + 1 + 0.25 thousand positions x 193,248 MACs)."""
    cfg = json.load(open(os.path.join(ROOT, "configs", "llicti_A.json")))
    data = tmp_path / "data"
    data.mkdir()
    cfg.update({"mode": mode, "test_data": str(data), "valid_data": str(data)})
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "main.py"), str(cfg_path)], cwd=tmp_path, env=dict(os.environ, PYTHONPATH=ROOT),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    exp = os.path.join("experiments", cfg["multi_exp_name"], "exp_0")
    log = (tmp_path / exp / "logs" / "exp_debug.log").read_text()
    if mode == "test":
        assert "No checkpoint exists" in log or needle in log
        return
    assert needle in log, log[-1500:]
    if mode == "flops_est":
        positions = sum((512 >> (s + 1)) ** 2 for s in range(5))
        assert "{:.3f} GMac".format(positions * 193248 / 1e9) in log, log[-1500:]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_main_eval_model_two_ranks(tmp_path):
    """eval_model under torchrun: the test images are sharded over the ranks (contiguous blocks), every image is coded and
    checked once, and the rate table -- the mean over all images, reduced over NCCL -- is the single-process table."""
    import re
    import socket
    from PIL import Image
    ocfg = O.OracleConfig()
    cfg = json.load(open(os.path.join(ROOT, "configs", "llicti_A.json")))
    data = tmp_path / "data"
    data.mkdir()
    for i, (h, w) in enumerate([(64, 96), (53, 77), (80, 64), (33, 47), (96, 96)]):
        Image.fromarray(O.synthetic_image(h, w, 40 + i).transpose(1, 2, 0)).save(data / f"img{i}.png")
    cfg["test_data"] = str(data)
    cfg_path = tmp_path / "cfg.json"
    cfg_path.write_text(json.dumps(cfg))
    exp = os.path.join("experiments", cfg["multi_exp_name"], "exp_0")
    sd = {k: torch.from_numpy(v) for k, v in O.synthetic_state_dict(ocfg).items()}
    totals = []
    for world in (1, 2):
        run = tmp_path / f"w{world}"
        (run / exp / "checkpoints").mkdir(parents=True)
        torch.save({"epoch": 1, "iteration": 2, "best_valid_loss": np.float64(1.0), "state_dict": sd}, run / exp / "checkpoints" / "model_best.pth.tar")
        cmd = [sys.executable, os.path.join(ROOT, "main.py"), str(cfg_path)]
        if world > 1:
            with socket.socket() as s:
                s.bind(("127.0.0.1", 0))
                port = s.getsockname()[1]
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                   "--master-port", str(port), os.path.join(ROOT, "main.py"), str(cfg_path)]
        r = subprocess.run(cmd, cwd=run, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        log = (run / exp / "logs" / "exp_debug.log").read_text()
        assert log.count("(Check: Decoded img matches original)") == 5, log[-3000:]
        assert log.count("Test Epoch:") == 1, log[-3000:]
        totals.append(re.findall(r"\(\(([0-9.]+)\)\)", log)[-1])
    assert totals[0] == totals[1], totals
