#!/usr/bin/env python
"""Does the CPU oracle reproduce itself (compress -> decompress) inside a process that has done what bench.py's
B200 arm does first -- CUDA initialised, libllicti_b200 loaded and used, pinned host buffers, an nvidia-smi
sampler thread, torch CPU work at the default thread count?  (Round 1: `bench.py --gpus 1` lost its cpu_baseline
leg to exactly that on one 32-thread host.)  Prints one line per repetition; exits 1 if any failed.

    python tools/oracle_selfcheck.py [--reps 3] [--size 512x768] [--config llicti_B.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--size", default="512x768")
    ap.add_argument("--config", default="llicti_B.json")
    args = ap.parse_args()
    H, W = (int(v) for v in args.size.split("x"))
    cfg = json.load(open(os.path.join(ROOT, "configs", args.config)))
    print(f"cpu_count {os.cpu_count()}, torch threads {torch.get_num_threads()}, cuda {torch.cuda.is_available()}", flush=True)
    if torch.cuda.is_available():                       # the state of bench.py's B200 arm
        from bench import ClockSampler
        from llicti_b200 import Codec, CodecConfig, synth, _lib as L
        ccfg = CodecConfig.from_json_dict(cfg, sub_len=0, numerics=L.NUM_TORCH_CUDA)
        codec = Codec(ccfg, synth.synthetic_state_dict(ccfg.chs))
        cs = ClockSampler(0)
        cs.start()
        rgb = torch.from_numpy(np.stack([synth.synthetic_image(H, W, 1000 + i) for i in range(4)])).pin_memory()
        for _ in range(3):
            bsl = codec.compress_images(rgb.numpy())
            assert np.array_equal(codec.decompress_images(bsl), rgb.numpy())
        print("clocks", cs.stop(), flush=True)
    from oracle import llicti_oracle as O
    ocfg = O.OracleConfig.from_dict(cfg)
    oc = O.OracleCodec(ocfg, O.synthetic_state_dict(ocfg))
    bad = 0
    for nt in (os.cpu_count() or 1, 16, 32, 64, 8):
        torch.set_num_threads(nt)
        for r in range(args.reps):
            img = O.synthetic_image(H, W, r)
            t0 = time.perf_counter()
            ok = np.array_equal(oc.decompress(oc.compress(img)), img)
            msg = "ok" if ok else "FAILED: " + O.diagnose_round_trip(oc, img)
            bad += not ok
            print(f"threads {nt:3d} image {r}: {msg} ({time.perf_counter() - t0:.1f} s)", flush=True)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
