#!/usr/bin/env python
"""oracle/make_ref.py -- TEST INFRASTRUCTURE.  Recipe that makes the UNMODIFIED reference runnable on the GPU box's
host cores: copies the reference's Python sources from where they lie (/root/reference, read-only, build container
only) into oracle/_ref/ (git-ignored: reference sources never enter the history; NOT gpurun-ignored: the directory
travels to the GPU box like a built .so) next to the three stand-ins for its uninstallable third-party imports
(oracle/refshims: compressai, torchac -> oracle/torchac_port.c, easydict).

    python oracle/make_ref.py            # called by __graft_entry__.build() when /root/reference exists

bench.py --impl reference (and its cpu_baseline leg) then times the reference's own LLICTI.compress / LLICTI.decompres
exactly as LLICTIAgent.eval_model calls them (agents/llicti_agent.py:135-149), cpu_baseline.kind = "reference".
Without oracle/_ref it falls back to the oracle port (kind = "port").
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DST = os.path.join(HERE, "_ref")


def make(force=False):
    if not os.path.isdir(REF):
        return None
    stamp = os.path.join(DST, ".made")
    if os.path.exists(stamp) and not force:
        shims = os.path.join(DST, "shims")                 # the stand-ins are this repo's: always the current ones
        shutil.rmtree(shims, ignore_errors=True)
        shutil.copytree(os.path.join(HERE, "refshims"), shims, ignore=shutil.ignore_patterns("__pycache__"))
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    keep = ("agents", "graphs", "loggers", "utils", "dataloaders", "configs")
    os.makedirs(DST)
    for d in keep:
        shutil.copytree(os.path.join(REF, d), os.path.join(DST, "reference", d),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    shutil.copy(os.path.join(REF, "main.py"), os.path.join(DST, "reference", "main.py"))
    shutil.copytree(os.path.join(HERE, "refshims"), os.path.join(DST, "shims"), ignore=shutil.ignore_patterns("__pycache__"))
    with open(stamp, "w") as f:
        f.write("copied from /root/reference by oracle/make_ref.py; not tracked by git\n")
    return DST


def load_reference_model(cfg_name, state_dict):
    """The reference's own LLICTI (graphs/models/LLICTI_nets.py) from oracle/_ref with `state_dict` loaded; None if
    oracle/_ref is absent.  cfg_name: a file of the reference's configs/ directory."""
    ref = os.path.join(DST, "reference")
    if not os.path.isdir(ref):
        return None
    import json
    import torch
    for p in (os.path.join(DST, "shims"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.dont_write_bytecode = True
    from easydict import EasyDict                      # shim
    from graphs.models.LLICTI_nets import LLICTI       # the reference
    with open(os.path.join(ref, "configs", cfg_name)) as f:
        cfg = EasyDict(json.load(f))
    model = LLICTI(cfg).eval()
    missing, unexpected = model.load_state_dict({k: torch.as_tensor(v) for k, v in state_dict.items()}, strict=False)
    assert not unexpected and all("conditional_prob_model" in k for k in missing), (missing, unexpected)
    return model


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
