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
