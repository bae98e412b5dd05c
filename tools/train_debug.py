"""Per-tensor, per-sub-network comparison of llicti_backward_dev with the golden reference gradients (debug aid)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import load_golden, oracle_config_for
from oracle import llicti_oracle as O
from llicti_b200 import _lib as L
from llicti_b200.codec import Codec, CodecConfig

name = sys.argv[1] if len(sys.argv) > 1 else "train_b_3x24x40"
g = load_golden(name)
ocfg = oracle_config_for(name)
sd = O.jittered_state_dict(ocfg, seed=1337)
codec = Codec(CodecConfig(num_scales=len(ocfg.dwtlevels), chs=ocfg.chs, numerics=L.NUM_TORCH_CPU, cnn_impl=L.CNN_FP32), sd)
rgb = torch.from_numpy(g["rgb"]).cuda()
sinfo = codec.forward_dev(rgb)
gs = [torch.full_like(s, 3.0 / rgb.numel()) for s in sinfo]
names = [k[5:] for k in g.files if k.startswith("grad/")]
grads = codec.backward_dev(rgb, gs, names)
G = ocfg.chs
for k in names:
    r = g["grad/" + k]
    q = grads[k].cpu().numpy().reshape(r.shape)
    rows = r.shape[0]
    nb = 4
    per = rows // nb
    out = []
    for b in range(nb):
        rr, qq = r[b * per:(b + 1) * per], q[b * per:(b + 1) * per]
        sc = np.abs(rr).max() + 1e-30
        i = np.unravel_index(np.abs(qq - rr).argmax(), rr.shape)
        out.append(f"{np.abs(qq - rr).max() / sc:.1e}(max {sc:.1e}; worst at {i}: {qq[i]:.4e} vs {rr[i]:.4e})")
    print(k.split("band.0.")[1], " | ".join(out))
