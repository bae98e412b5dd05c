#!/usr/bin/env python
"""Benchmark of the LLICTI compress/decompress hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c0..c4] [--impl b200|reference]

A step = one pass of the hot path (compress then decompres) over ONE batch of synthetic images.
Workloads are BASELINE.json's configs:

  c0  configs[0]  llicti_A, 24 x 768x512, torchac-compatible streams (the reference's own CPU-runnable case)
  c1  configs[1]  llicti_B, the same 24 images, torchac-compatible byte-exact mode
  c2  configs[2]  llicti_A, 100 x 2040x1356 as one batch per step, interleaved-substream coder   <- default at N = 1
  c3  configs[3]  llicti_A, 512 x 3840x2160 sharded over the N GPUs (512 / N distinct images per rank, batches of 32)
                                                                                                <- default at N > 1
  c4  configs[4]  llicti_A, 50,000 x 512x512 sharded over the N GPUs (batches of 512)

Every rank generates ITS OWN images (seeded by global image index); step k codes batch k mod (resident batches).
Rank 0 prints ONE JSON line on stdout:

  value         whole-job round-trip throughput (MP/s), batches resident in HBM, max over ranks of the device time
  e2e           the same through llicti_encode_host / llicti_decode_host with pinned HOST buffers
  roofline      dominant kernel class: algorithmic bytes (or flops) per launch / its CUDA-event time
  per_config    short passes over the other configs (N = 1: c0, c1, c3, c4; N > 1: c4), same code path
  cpu_baseline  the CPU oracle on a bounded sample, run in a fresh process (`bench.py --impl reference`); a failure
                there is recorded in the JSON line and never discards the GPU measurement

`--impl reference` times the reference's CPU path alone: the UNMODIFIED reference sources from oracle/_ref (copied
there by oracle/make_ref.py in the build container; the directory travels to the GPU box) when present, else
oracle/llicti_oracle.py, which restates it with the same cost structure.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENCODE_ONLY = bool(os.environ.get("LLICTI_BENCH_ENCODE_ONLY"))   # debugging aid: time the encoder alone

WORKLOADS = {
    # name: description, config, images per batch, H, W, sub_len, images of the whole job (sharded over the ranks)
    "c0": dict(desc="configs[0]: llicti_A eval_model, 24 synthetic 768x512 RGB images, torchac-compatible streams",
               cfg="llicti_A.json", batch=24, H=512, W=768, sub_len=0, total=24, shard=False),
    "c1": dict(desc="configs[1]: llicti_B eval_model, 24 synthetic 768x512 RGB images, torchac-compatible byte-exact mode",
               cfg="llicti_B.json", batch=24, H=512, W=768, sub_len=0, total=24, shard=False),
    # (one batch = the config's 100 images: the decoder's chain-parallel kernels fill the machine better with 2 x 17.6 k chains
    #  per launch than with one -- 1419 against 1333 MP/s in two batches of 50, profiles/r02/bench_c2_batch100.json)
    "c2": dict(desc="configs[2]: llicti_A, 100 synthetic 2040x1356 images (one batch per step), interleaved-substream coder",
               cfg="llicti_A.json", batch=100, H=1356, W=2040, sub_len=2048, total=100, shard=False),
    "c3": dict(desc="configs[3]: llicti_A, 512 synthetic 3840x2160 images sharded over the GPUs, batches of 32, "
                    "interleaved-substream coder",
               cfg="llicti_A.json", batch=32, H=2160, W=3840, sub_len=2048, total=512, shard=True),
    "c4": dict(desc="configs[4]: llicti_A, 50,000 synthetic 512x512 images sharded over the GPUs, batches of 512, "
                    "interleaved-substream coder",
               cfg="llicti_A.json", batch=512, H=512, W=512, sub_len=2048, total=50000, shard=True),
}
METRIC = "encode+decode round-trip megapixels/s (compress then decompres of every image)"
MAC_PER_POS = {88: (53152, 61600, 78496), 60: (29520, 35280, 46800)}
# MACs the tcgen05 kernel EXECUTES per position and band: layer-0 depth padded to chunks x 4 taps x 8 slots (64 / 96 / 160),
# widths padded 88 -> 96 / 60 -> 64 (two bias slots, multiples of 16), layer 2 padded 15 -> 16 outputs; all four sub-networks
MAC_EXEC_PER_POS = {88: tuple(4 * (96 * k0 + 96 * 96 + 16 * 96) for k0 in (64, 96, 160)),
                    60: tuple(4 * (64 * k0 + 64 * 64 + 16 * 64) for k0 in (64, 96, 160))}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def load_cfg(name):
    with open(os.path.join(ROOT, "configs", name)) as f:
        return json.load(f)


# =============================================================================================
# reference arm / cpu_baseline: the CPU oracle, one image per step
# =============================================================================================
def cpu_oracle_pass(cfg_name, cfg_json, H, W, steps, warmup, seed0=0):
    """Time the reference's CPU path (compress + decompres of one image per step): the UNMODIFIED reference from
    oracle/_ref when it travelled with the repo (kind "reference": its own LLICTI.compress / LLICTI.decompres, called as
    LLICTIAgent.eval_model calls them), else the oracle's restatement (kind "port").  Every image must round-trip; a
    mismatch is diagnosed (which stream's table differs between encode and decode) before raising."""
    from oracle import llicti_oracle as O
    from oracle import make_ref
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.OracleConfig.from_dict(cfg_json)
    sd = O.synthetic_state_dict(ocfg)
    codec = O.OracleCodec(ocfg, sd, sub_len=0)
    model = None if os.environ.get("LLICTI_BENCH_PORT") else make_ref.load_reference_model(cfg_name, sd)
    kind = "reference" if model is not None else "port"
    times = []
    for it in range(warmup + steps):
        img = O.synthetic_image(H, W, seed0 + it)
        if model is not None:
            x = torch.from_numpy(img.astype(np.float32) / np.float32(255.0))[None]      # what the reference's ToTensor() loader yields
            with torch.no_grad():
                t0 = time.perf_counter()
                bsl, _ = model.compress(x.clone())
                t1 = time.perf_counter()
                rec_f = model.decompres(bsl, torch.device("cpu"))
                t2 = time.perf_counter()
            ok = ((x - rec_f) * 255).abs().max().item() < 0.5                             # eval_model's own check (llicti_agent.py:151)
        else:
            t0 = time.perf_counter()
            bsl = codec.compress(img)
            t1 = time.perf_counter()
            rec = codec.decompress(bsl)
            t2 = time.perf_counter()
            ok = np.array_equal(rec, img)
        if not ok:
            raise RuntimeError(f"CPU {kind} round trip of image {seed0 + it} is not lossless: " + O.diagnose_round_trip(codec, img))
        log(f"[cpu {kind}] step {it}: enc {t1 - t0:.2f}s dec {t2 - t1:.2f}s")
        if it >= warmup:
            times.append((t1 - t0, t2 - t1))
    enc = sum(t[0] for t in times)
    dec = sum(t[1] for t in times)
    px = H * W * len(times) / 1e6
    return {"value": px / (enc + dec), "encode_mpps": px / enc, "decode_mpps": px / dec,
            "ms_per_step": 1e3 * (enc + dec) / len(times), "cores": torch.get_num_threads(), "kind": kind}


def reference_sample_shape(wl):
    """The reference materialises dense H x W x Lp CDF tables (5 fp32 temporaries of that size): a 2040x1356 image takes
    minutes and ~10 GB, a 4K image cannot be held at all.  Its cost per pixel does not depend on the image size, so
    big shapes are sampled as 768x512 images of the same generator."""
    return (wl["H"], wl["W"]) if wl["H"] * wl["W"] <= 768 * 512 else (512, 768)


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cfg = load_cfg(wl["cfg"])
    sh, sw = reference_sample_shape(wl)
    r = cpu_oracle_pass(wl["cfg"], cfg, sh, sw, args.steps, args.warmup)
    sample = (f"{args.steps} step(s) of 1 synthetic {sw}x{sh} image each, compress+decompres, {r['cores']} torch threads, "
              f"model config {wl['cfg']}; " + ("the unmodified reference (oracle/_ref: its sources + stand-ins for compressai / "
              "torchac / easydict)" if r["kind"] == "reference" else "the oracle's restatement of the reference (oracle/_ref absent)"))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "model_config": wl["cfg"], "sample": sample},
            "encode_mpps": r["encode_mpps"], "decode_mpps": r["decode_mpps"],
            "cpu_baseline": {"value": r["value"], "unit": "MP/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_train_pass(batch=4, patch=160, steps=3):
    """The reference's training step on the host cores (bounded sample: `batch` patches per step): the UNMODIFIED
    reference from oracle/_ref -- model.train(); model(x); TrainRLossList; loss.backward() as agents/llicti_agent.py:52-61 --
    else the oracle's autograd restatement (kind "port").  Prints one JSON line."""
    from oracle import llicti_oracle as O
    from oracle import make_ref
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.OracleConfig()
    sd = O.synthetic_state_dict(ocfg)
    model = None if os.environ.get("LLICTI_BENCH_PORT") else make_ref.load_reference_model("llicti_A.json", sd)
    kind = "reference" if model is not None else "port"
    if model is not None:
        from graphs.losses.rate_dist import TrainRLossList         # the reference's (oracle/_ref is on sys.path now)
        model.train()
        loss_fn = TrainRLossList()
    times = []
    for it in range(1 + steps):
        rgb = np.stack([O.synthetic_image(patch, patch, 4000 + batch * it + i) for i in range(batch)])
        t0 = time.perf_counter()
        if model is not None:
            x = torch.from_numpy(rgb.astype(np.float32) / np.float32(255.0))
            model.zero_grad()
            loss, _ = loss_fn.forward(torch.numel(x), model(x))
            loss.backward()
        else:
            O.train_loss_and_grads(ocfg, sd, rgb)
        if it >= 1:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    print(json.dumps({"value": batch / t, "unit": "patches/s", "ms_per_step": 1e3 * t, "cores": torch.get_num_threads(), "kind": kind,
                      "sample": f"{steps} step(s) of {batch} synthetic {patch}x{patch} patches, forward + backward, llicti_A, "
                                f"{torch.get_num_threads()} torch threads"}), flush=True)


def cpu_train_leg():
    """cpu_train_pass in a fresh process without CUDA; never fatal."""
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-train-leg"], capture_output=True, text=True, timeout=300,
                           env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), preexec_fn=_unbound)
        line = next((ln for ln in reversed(r.stdout.splitlines()) if ln.startswith("{")), None)
        if r.returncode != 0 or line is None:
            return {"error": f"cpu training process exited {r.returncode}: {r.stderr.strip().splitlines()[-1:] or ''}"}
        return json.loads(line)
    except Exception as e:       # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def cpu_baseline_leg(workload, steps=4):
    """The CPU oracle on a bounded sample, in a FRESH process (the state the reference arm runs in): whatever this
    process has loaded (CUDA, NCCL, the codec library, pinned allocations) cannot disturb it, and a failure becomes
    an "error" entry of the JSON line instead of discarding the GPU measurement."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", workload, "--steps", str(steps),
           "--warmup", "0"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), preexec_fn=_unbound)
        line = next((ln for ln in reversed(r.stdout.splitlines()) if ln.startswith("{")), None)
        if r.returncode != 0 or line is None:
            return {"error": f"cpu oracle process exited {r.returncode}: {r.stderr.strip().splitlines()[-1:] or ''}", "kind": "port"}
        d = json.loads(line)
        cb = d["cpu_baseline"]
        cb.update({"encode_mpps": d["encode_mpps"], "decode_mpps": d["decode_mpps"]})
        return cb
    except Exception as e:       # noqa: BLE001 -- a diagnostic leg must never take the measurement down
        return {"error": f"{type(e).__name__}: {e}", "kind": "port"}


# =============================================================================================
# B200 arm
# =============================================================================================
# How each kernel class is bounded: the CNN by the tensor pipe, everything else is
# byte/integer work whose ceiling is HBM bandwidth -- the serial coder chains and the erfc-heavy
# CDF kernels sit far below it by nature (latency / instruction issue), which the fractions show.
BOUND = {"cnn": "tensor"}


def windowed_bands(geom, n, sub_len):
    """Which (scale, band) pairs decode through the windowed schedules (window rows through HBM, then the chain kernel): all
    of them for torchac-compatible streams; for the substream container the bands with too few chains for the lane decoder
    (the library's rule, kernels_decode.cu: ceil(n * S / 10) warps of chains >= LLICTI_LANE_MIN_WARPS, default 148)."""
    S = geom.num_scales
    if sub_len <= 0:
        return {(s, b) for s in range(S) for b in range(3)}
    min_warps = int(os.environ.get("LLICTI_LANE_MIN_WARPS", "148"))
    lanes = os.environ.get("LLICTI_DECODE_LANES", "1") != "0"
    return {(s, b) for s in range(S) for b in range(3)
            if lanes and -(-n * geom.num_sub[s][b] // 10) < min_warps}


def algorithmic_work(geom, chs, n, blob_bytes, sub_len):
    """Per-step ALGORITHMIC bytes / flops of each kernel class for n images:
    what the stage has to read and write by definition, not what a particular schedule moves internally."""
    S = geom.num_scales
    pos = [geom.Hs[s] * geom.Ws[s] for s in range(S)]
    sym_band = [[geom.crop_h[s][b] * geom.crop_w[s][b] for b in range(3)] for s in range(S)]
    macs = sum(pos[s] * sum(MAC_PER_POS[chs]) for s in range(S))
    coded_pos = sum(sum(sym_band[s]) for s in range(S))
    win = windowed_bands(geom, n, sub_len)
    win_pos = sum(sym_band[s][b] for (s, b) in win)
    return {
        "split": n * (3 * geom.H * geom.W + 2 * 12 * sum(pos)),                 # u8 in, int16 planes out
        "cnn_flops": 2 * n * 2.0 * macs,                                        # the CNN runs in both directions
        "cnn": 2 * n * sum(pos[s] * (2 * 3 * (b + 1) + 240) for s in range(S) for b in range(3)),
        "bounds": n * coded_pos * (240 + 6 + 12),                               # 258 B per (position, band)
        "encode": n * geom.symbols * 4 + blob_bytes,                            # 4 B bounds in + bytes out
        # decode side, 246 B per (position, band) = params + symbols, plus the stream bytes: the windowed bands' share belongs
        # to the window kernel (the rows it writes and the chain kernel reads back are internal), the rest to the class that
        # decodes straight from the params (lane / group decoder)
        # (torchac-compatible streams: the wavefront / piped schedules produce and consume inside ONE kernel class, "decode")
        "window": n * win_pos * (240 + 6),
        "decode": n * (coded_pos if sub_len <= 0 else coded_pos - win_pos) * (240 + 6) + blob_bytes,
        "merge": n * (2 * 12 * sum(pos) + 3 * geom.H * geom.W),
    }


def internal_traffic(geom, n, sub_len):
    """Bytes the windowed decode schedules move through HBM on top of the algorithmic ones: 3 x 64 B of window rows per
    (position, band) written by the producers and read back by the chains."""
    win_pos = sum(geom.crop_h[s][b] * geom.crop_w[s][b] for (s, b) in windowed_bands(geom, n, sub_len))
    return {"window_rows_written": n * win_pos * 3 * 64, "window_rows_read": n * win_pos * 3 * 64}


class Workload:
    """One config on this rank: its codec, its resident batches and the timed loops over them."""

    def __init__(self, name, args, rank, world, local_rank, max_batches):
        from llicti_b200 import Codec, CodecConfig, _lib as L
        from llicti_b200 import synth
        from llicti_b200.shard import shard_range
        self.name, self.wl, self.args = name, WORKLOADS[name], args
        wl = self.wl
        self.rank, self.world = rank, world
        self.cfg = load_cfg(wl["cfg"])
        self.dev = torch.device("cuda", local_rank)
        self.L = L
        self.ccfg = CodecConfig.from_json_dict(self.cfg, sub_len=wl["sub_len"], numerics=L.NUM_TORCH_CUDA, cnn_impl=args.cnn,
                                               device=local_rank, decode_impl=args.decode_impl)
        self.sd = synth.synthetic_state_dict(self.ccfg.chs, self.ccfg.num_mixtures, int(self.cfg["Evens"][0]), int(self.cfg["Odds"][0]))
        if args.weights_npz:                      # e.g. tests/golden/ckpt_A_trained.npz: weights short-trained by the reference's own code
            with np.load(args.weights_npz) as z:
                self.sd = {k: z[k] for k in z.files}
        self.codec = Codec(self.ccfg, self.sd)
        H, W = wl["H"], wl["W"]
        self.H, self.W = H, W
        self.geom = self.codec.geometry(H, W)
        self.st = 2 ** self.geom.num_scales
        # the job's images: sharded configs give every rank its contiguous block of distinct images; the single-GPU
        # configs (c0-c2) are replicated per rank with rank-specific seeds (distinct images as well)
        total = args.images * world if args.images and wl["shard"] else (args.images or wl["total"])
        if wl["shard"]:
            first, last = shard_range(total, rank, world)
        else:
            first, last = rank * total, (rank + 1) * total
        self.job_images, self.shard_images = (total if wl["shard"] else total * world), last - first
        self.n = min(args.batch or wl["batch"], self.shard_images)
        self.job_batches = -(-self.shard_images // self.n)                    # batches this rank would code for the whole job
        nb = max(1, min(self.job_batches, max_batches))
        t0 = time.perf_counter()
        self.batches = []
        for b in range(nb):
            k0 = first + b * self.n
            cnt = min(self.n, last - k0)
            if cnt < self.n:                      # ragged tail of the shard: wrap around to keep the batch shape
                k0 = last - self.n
            self.batches.append(synth.synthetic_batch_torch(self.n, H, W, 1000 + k0, self.dev))
        torch.cuda.synchronize()
        self.gen_s = time.perf_counter() - t0
        self.x00 = [b[:, :, ::self.st, ::self.st].contiguous() for b in self.batches]
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)      # > 126 MB L2
        self.enc_out = None
        self.rec_out = None

    def close(self):
        self.codec.close()
        del self.batches, self.x00, self.flush, self.enc_out, self.rec_out
        torch.cuda.empty_cache()

    def barrier(self):
        if self.world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def dev_step(self, k):
        rgb_d, x00_d = self.batches[k % len(self.batches)], self.x00[k % len(self.batches)]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self.flush.zero_()
        ev[0].record()
        self.enc_out = self.codec.encode_dev(rgb_d, self.enc_out)
        blob, off, mm = self.enc_out
        ev[1].record()
        self.flush.zero_()                      # decode starts cold as well (outside both timed spans)
        ev[2].record()
        if ENCODE_ONLY:                         # kernel experiments whose streams are not decodable (tools/)
            ev[3].record()
            torch.cuda.synchronize()
            return ev[0].elapsed_time(ev[1]), 1e-3, off, rgb_d, rgb_d
        self.rec_out = self.codec.decode_dev(blob, off, mm, x00_d, self.n, self.H, self.W, self.rec_out)
        ev[3].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), off, self.rec_out, rgb_d

    def measure_device(self, steps, warmup, profile=True):
        """W warm-up steps, then K timed steps with the batches resident in HBM; returns the raw per-rank numbers."""
        codec = self.codec
        for k in range(max(warmup, 1)):
            _, _, off, rec, rgb_d = self.dev_step(k)
            assert torch.equal(rec, rgb_d), f"{self.name}: device round trip is not lossless"
        codec.check_status()
        self.barrier()
        launches0 = codec.launches
        codec.decode_stats()
        t_enc = t_dec = 0.0
        blob_bytes = 0
        per_step = []
        for k in range(steps):
            a, b, off, rec, rgb_d = self.dev_step(warmup + k)
            t_enc += a
            t_dec += b
            per_step.append(a + b)
            blob_bytes += int(off[-1].item())
        ok = torch.equal(rec, rgb_d)
        self.barrier()
        assert ok, f"{self.name}: device round trip is not lossless"
        out = {"t_enc": t_enc, "t_dec": t_dec, "blob_bytes": blob_bytes / steps, "launches": codec.launches - launches0,
               "dstats": codec.decode_stats(), "per_step_ms": per_step}
        if profile:
            # the same K steps once more with the library's per-kernel-class events (they need eager launches: the
            # decode of the timed region above may be replayed as one CUDA graph)
            codec.profile(True)
            p_enc = p_dec = 0.0
            for k in range(steps):
                a, b, *_ = self.dev_step(warmup + k)
                p_enc += a
                p_dec += b
            out["prof"] = codec.profile_read()
            out["p_enc"], out["p_dec"] = p_enc, p_dec
            codec.profile(False)
            codec.decode_stats()
            self.barrier()
        return out

    def measure_host(self, steps, warmup):
        """The same K steps through the *_host entry points: pinned host buffers, host<->device copies inside the timed
        region.  Host copies exist for at most 4 of the resident batches (pinned memory is not free)."""
        if ENCODE_ONLY:
            return {"h_enc": 1e-3 * steps, "h_dec": 1e-3 * steps, "h2d": 0, "d2h": 0}
        codec, n, H, W = self.codec, self.n, self.H, self.W
        nb = min(len(self.batches), 4)
        rgb_h = [self.batches[b].cpu().pin_memory().numpy() for b in range(nb)]
        x00_h = [np.ascontiguousarray(r[:, :, ::self.st, ::self.st]) for r in rgb_h]
        out_pinned = torch.empty(n * int(self.geom.max_stream_bytes), dtype=torch.uint8).pin_memory().numpy()
        rec_pinned = torch.empty((n, 3, H, W), dtype=torch.uint8).pin_memory().numpy()

        def host_step(k):
            rgb_np, x00_np = rgb_h[k % nb], x00_h[k % nb]
            self.flush.zero_()
            torch.cuda.synchronize()
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
            blob, off, mm = codec.encode_host(rgb_np, out_pinned)
            e1.record()
            self.flush.zero_()
            e2.record()
            rec = codec.decode_host(blob, off, mm, x00_np, n, H, W, rec_pinned)
            e3.record()
            torch.cuda.synchronize()
            h2d = rgb_np.nbytes + blob.nbytes + off.nbytes + mm.nbytes + x00_np.nbytes
            d2h = blob.nbytes + off.nbytes + mm.nbytes + rec.nbytes
            return e0.elapsed_time(e1), e2.elapsed_time(e3), h2d, d2h, rec, rgb_np

        for k in range(max(1, min(warmup, 2))):
            host_step(k)
        self.barrier()
        h_enc = h_dec = 0.0
        for k in range(steps):
            a, b, h2d, d2h, rec, rgb_np = host_step(k)
            h_enc += a
            h_dec += b
        ok = np.array_equal(rec, rgb_np)
        self.barrier()
        assert ok, f"{self.name}: host round trip is not lossless"
        return {"h_enc": h_enc, "h_dec": h_dec, "h2d": h2d, "d2h": d2h}

    def bpp_delta(self):
        """Rate of the substream container against the torchac-compatible streams (= the reference's bytes when the
        network outputs agree) of the same images, untimed, on a few images of the first batch."""
        from llicti_b200 import Codec, CodecConfig
        if self.wl["sub_len"] <= 0 or ENCODE_ONLY:
            return None
        m = min(self.n, 4)
        ns = 9 * self.geom.num_scales
        sub = self.codec.encode_dev(self.batches[0][:m].contiguous())
        sub_bytes = int(sub[1][m * ns].item()) + 5 * m          # + the container's mode tag in the header row
        compat = Codec(CodecConfig.from_json_dict(self.cfg, sub_len=0, numerics=self.L.NUM_TORCH_CUDA, cnn_impl=self.args.cnn,
                                                  device=self.dev.index, decode_impl=self.args.decode_impl), self.sd)
        _, c_off, _ = compat.encode_dev(self.batches[0][:m].contiguous())
        compat_bytes = int(c_off[-1].item())
        compat.close()
        del compat
        torch.cuda.empty_cache()
        return {"value": (sub_bytes - compat_bytes) / compat_bytes, "images": m, "substream_bytes": sub_bytes,
                "torchac_compatible_bytes": compat_bytes,
                "note": "(substream container - torchac-compatible streams) / torchac-compatible streams, all length "
                        "tables and flush bytes counted"}


def _fp32_roofline(flop_per_step, cnn_ms_per_step, dev):
    """The training kernels are fp32 CUDA-core kernels: their ceiling is the FMA pipe (SMs x 128 lanes x 2 FLOP x the
    maximum SM clock), not the tensor pipe and not HBM."""
    if not cnn_ms_per_step:
        return None
    p = torch.cuda.get_device_properties(dev)
    try:
        import pynvml
        pynvml.nvmlInit()
        mhz = pynvml.nvmlDeviceGetMaxClockInfo(pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0), pynvml.NVML_CLOCK_SM)
    except Exception:       # noqa: BLE001
        mhz = p.clock_rate / 1e3 if getattr(p, "clock_rate", 0) else 1965.0
    peak = p.multi_processor_count * 128 * 2 * mhz * 1e6 / 1e12
    ach = flop_per_step / (cnn_ms_per_step * 1e-3) / 1e12
    return {"bound": "fp32 FMA pipe", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "kernels": "cnn_forward_train_kernel + cnn_backward_kernel (algorithmic FLOP: one forward, the recompute of two layers, five backward products)"}


def train_step_pass(local_rank, steps, warmup, batch=32, patch=160):
    """One training step = llicti_set_weights_dev + llicti_train_forward_dev + llicti_backward_dev on a batch resident in HBM
    (the optimizer is torch's and outside the timed region, as it is outside the library)."""
    from llicti_b200 import _lib as L
    from llicti_b200.codec import Codec, CodecConfig, PREFIX
    from llicti_b200 import synth
    sd = synth.synthetic_state_dict()
    codec = Codec(CodecConfig(cnn_impl=L.CNN_FP32, device=local_rank), sd)
    dev = codec.device
    wts = {k: torch.from_numpy(v).to(dev) for k, v in sd.items() if k.startswith(PREFIX) and "conditional_prob_model" not in k}
    names = list(wts)
    rgbs = [synth.synthetic_batch_torch(batch, patch, patch, 4000 + 64 * k, dev) for k in range(2)]
    numel = batch * 3 * patch * patch
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms, loss = [], None
    for k in range(warmup + steps):
        if k == warmup:
            codec.profile(True)
        rgb = rgbs[k % 2]
        torch.cuda.synchronize(dev)
        ev[0].record()
        codec.set_weights_dev(wts)
        sinfo, kept = codec.train_forward_dev(rgb)
        gs = [torch.full_like(t, 3.0 / numel) for t in sinfo]
        grads = codec.backward_dev(rgb, gs, names, kept=kept)
        ev[1].record()
        torch.cuda.synchronize(dev)
        if k >= warmup:
            ms.append(ev[0].elapsed_time(ev[1]))
        loss = sum(float(t.double().sum()) for t in sinfo) / numel * 3
        assert all(bool(torch.isfinite(g).all()) for g in grads.values())
    prof = codec.profile_read()
    codec.profile(False)
    codec.close()
    t = float(np.mean(ms))
    G = 88
    pos = sum(batch * (patch >> (s + 1)) ** 2 for s in range(5))
    fwd = sum(2 * 4 * (k0 * G + G * G + 15 * G) for k0 in (48, 72, 120))                      # FLOP per position, three bands
    bwd = sum(2 * 4 * (2 * k0 * G + 3 * G * G + 30 * G) for k0 in (48, 72, 120))              # recompute of two layers + five products
    return {"workload": f"training step, llicti_A, {batch} patches of {patch}x{patch} (the reference's batch_size / patch_size), fp32",
            "ms_per_step": t, "patches_per_s": batch / t * 1e3, "value": batch * patch * patch / t / 1e3, "unit": "MP/s",
            "steps": steps, "warmup": warmup, "loss_bpp": loss,
            "kernel_ms_per_step": {k: v[0] / steps for k, v in prof.items() if v[1]},
            "cnn_class_tflops": pos * (fwd + bwd) / (prof["cnn"][0] / steps * 1e-3) / 1e12 if prof.get("cnn", (0, 0))[0] else None,
            "roofline": _fp32_roofline(pos * (fwd + bwd), prof.get("cnn", (0, 0))[0] / steps, dev),
            "note": "cnn class = fp32 forward + cnn_backward_kernel (which recomputes the two hidden layers); "
                    "bounds class = self_info_kernel + self_info_grad_kernel"}


def reduce_workload(w, dev_r, host_r, steps):
    """Max over ranks of the times, sum over ranks of the work; every rank returns the same dict."""
    from llicti_b200.shard import reduce_stats
    px = w.n * w.H * w.W
    sums = [px, dev_r["blob_bytes"], dev_r["launches"], host_r["h2d"] if host_r else 0, host_r["d2h"] if host_r else 0]
    maxs = [dev_r["t_enc"], dev_r["t_dec"], dev_r["t_enc"] + dev_r["t_dec"],
            host_r["h_enc"] if host_r else 0, host_r["h_dec"] if host_r else 0,
            (host_r["h_enc"] + host_r["h_dec"]) if host_r else 0, -dev_r["t_dec"], w.job_batches]
    (px_total, bytes_total, launches_total, h2d_total, d2h_total), mx = reduce_stats(sums, maxs, device=w.dev)
    t_enc, t_dec, t_rt, h_enc, h_dec, h_rt, neg_min_dec, job_batches = mx
    mp = px_total / 1e6
    K = steps
    res = {"value": mp * K / (t_rt / 1e3), "ms_per_step": t_rt / K, "encode_mpps": mp * K / (t_enc / 1e3),
           "decode_mpps": mp * K / (t_dec / 1e3), "encode_ms_per_step": t_enc / K, "decode_ms_per_step": t_dec / K,
           "decode_ms_per_step_rank_min_max": [-neg_min_dec / K, t_dec / K],
           "bpp": bytes_total * 8 / px_total, "bpsp": bytes_total * 8 / (px_total * 3), "compressed_bytes_per_step": bytes_total,
           "gpu_launches": int(launches_total), "px_per_step": px_total,
           # strong-scaling view of the sharded configs: the slowest rank's share of the whole job at the measured rate
           "job": {"images": w.job_images, "images_per_rank": w.shard_images, "batches_per_rank": int(job_batches),
                   "seconds_for_whole_job": job_batches * t_rt / K / 1e3}}
    if host_r:
        res["e2e"] = {"value": mp * K / (h_rt / 1e3), "unit": "MP/s", "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                      "encode_mpps": mp * K / (h_enc / 1e3), "decode_mpps": mp * K / (h_dec / 1e3),
                      "api": "llicti_encode_host + llicti_decode_host, pinned host buffers"}
    return res


def rooflines(w, dev_r, args, steps):
    """Per-class roofline entries from rank 0's per-kernel-class events."""
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        pk = json.load(open(pk_path))
        peaks = {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops_sustained": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                 "src": "measured (MEASURED_PEAKS.json; sustained bf16 figure: the kernels are timed inside a long step)"}
    prof, K = dev_r["prof"], steps
    work_step = algorithmic_work(w.geom, w.ccfg.chs, w.n, dev_r["blob_bytes"], w.wl["sub_len"])
    kernel_ms = {k: v[0] / K for k, v in prof.items()}
    step_ms = (dev_r["p_enc"] + dev_r["p_dec"]) / K
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic_db = json.load(open(tpath)).get(w.name, {})

    def roofline_of(cls):
        groups = max(prof[cls][1] / K, 1.0)
        ms = kernel_ms[cls] / groups
        if ms <= 0 or cls not in work_step:
            return None
        if BOUND.get(cls) == "tensor" and args.cnn == 1:
            ach = work_step["cnn_flops"] / groups / (ms * 1e-3) / 1e12
            r = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s"}
        else:
            ach = work_step[cls] / groups / (ms * 1e-3) / 1e9
            r = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
        r["frac"] = r["achieved"] / r["peak"]
        t = traffic_db.get(cls)      # dram bytes of one captured launch (ncu --set full), scaled by work units to the average launch
        r["traffic"] = (t["dram_bytes"] / t["units"] * (w.n * t["units_per_image_step"] / groups)) if t else None
        r.update({"kernel": cls, "launch_groups_per_step": groups, "ms_per_launch_group": ms,
                  "algorithmic_per_launch_group": (work_step["cnn_flops"] if r["bound"] == "tensor" else work_step[cls]) / groups,
                  "share_of_step": kernel_ms[cls] / step_ms})
        if t:
            r["traffic_source"] = t["source"]
        return r

    dom = max((c for c in kernel_ms if c in work_step), key=lambda c: kernel_ms[c])
    roof = roofline_of(dom)
    roof["peak_source"] = peaks["src"]
    allr = {c: {k: v for k, v in r.items() if k in ("bound", "achieved", "peak", "unit", "frac", "share_of_step")}
            for c in kernel_ms if kernel_ms[c] > 0 for r in [roofline_of(c)] if r}
    cnn_tflops = work_step["cnn_flops"] / (max(kernel_ms["cnn"], 1e-9) * 1e-3) / 1e12
    if roof.get("kernel") == "cnn" and args.cnn == 1:      # what the tensor pipe really multiplies (MMA shape padding included)
        pad = sum(MAC_EXEC_PER_POS[w.ccfg.chs]) / sum(MAC_PER_POS[w.ccfg.chs])
        roof["executed"] = {"achieved": roof["achieved"] * pad, "frac": roof["frac"] * pad, "unit": "TFLOP/s",
                            "note": "MACs the kernel issues (layer-0 depth and layer widths padded to MMA shapes) / the same time; "
                                    "`achieved` and `frac` count the reference's algorithmic MACs only"}
    return roof, allr, kernel_ms, cnn_tflops, internal_traffic(w.geom, w.n, w.wl["sub_len"])


ALL_CPUS = set(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None


def _unbound():
    """preexec_fn of the CPU legs: they run on every host core this process started with, not on the GPU's NUMA node."""
    if ALL_CPUS:
        os.sched_setaffinity(0, ALL_CPUS)


def run_b200(args, rank, world, local_rank):
    torch.cuda.set_device(local_rank)
    from llicti_b200.shard import bind_host_to_gpu
    bound = bind_host_to_gpu(local_rank)          # before any pinned allocation: staging buffers local to the GPU's PCIe root
    log(f"[rank {rank}] host threads bound to {len(bound)} CPUs local to GPU {local_rank}" if bound else f"[rank {rank}] no CPU binding")
    primary = args.workload or ("c2" if world == 1 else "c3")
    K, Wm = args.steps, args.warmup

    # ---- primary workload: full measurement -------------------------------------------------------
    w = Workload(primary, args, rank, world, local_rank, max_batches=Wm + K)
    clocks = ClockSampler(local_rank)
    w.barrier()
    clocks.start()
    dev_r = w.measure_device(K, Wm, profile=True)
    host_r = w.measure_host(K, Wm)
    clk = clocks.stop()          # sampled over both timed regions (device-resident and end-to-end)
    res = reduce_workload(w, dev_r, host_r, K)
    bpp_delta = w.bpp_delta() if rank == 0 else None
    roof = allr = kernel_ms = cnn_tflops = internal = None
    if rank == 0:
        roof, allr, kernel_ms, cnn_tflops, internal = rooflines(w, dev_r, args, K)
    wl = w.wl
    cfg_entry = {"workload": wl["desc"], "model_config": wl["cfg"], "images_per_step_per_gpu": w.n, "height": w.H, "width": w.W,
                 "sub_len": wl["sub_len"], "cnn_impl": "tcgen05" if args.cnn == 1 else "fp32-cuda-core",
                 "cnn_operands": {0: "fp32", 1: "bf16 (fp32 accumulate)", 2: "fp16 (fp32 accumulate)"}.get(int(w.codec.cnn_operands), "?"),
                 "decode_impl": "default" if args.decode_impl == 0 else "legacy-warp",
                 "distinct_batches_resident_per_gpu": len(w.batches),
                 "weights": (args.weights_npz or "llicti_b200.synth.synthetic_state_dict(seed=1337) (shipped checkpoint absent)"),
                 "images": f"llicti_b200.synth.synthetic_batch_torch, image k of the job seeded by 1000 + k; generated in {w.gen_s:.1f} s",
                 "l2": "256 MiB buffer written before every timed encode and decode (L2 flushed)",
                 "host_cpus_bound": (f"{len(bound)} CPUs local to the GPU (NVML affinity), set before the pinned buffers were allocated" if bound
                                     else "none (all of the process's CPUs are local to the GPU, or NVML gave no affinity)"),
                 "parallelism": (f"{w.job_images} images sharded over {world} GPU(s): {w.shard_images} distinct images per rank"
                                 if wl["shard"] else f"{w.shard_images} distinct images per rank on {world} GPU(s)") +
                                ", no data-path collective; NCCL only reduces the statistics"}
    dstats = {k: v / K for k, v in dev_r["dstats"].items()}
    p_enc, p_dec = dev_r["p_enc"], dev_r["p_dec"]
    w.close()

    # ---- the other configs: short passes, same code path (device-resident timing only) ------------------
    per_config = {}
    others = [] if args.no_per_config else ([c for c in ("c0", "c1", "c3", "c4") if c != primary] if world == 1 else
                                            [c for c in ("c4",) if c != primary])
    for name in others:
        try:
            ks, ws = min(K, 3), min(Wm, 2)
            o = Workload(name, args, rank, world, local_rank, max_batches=ks + ws)
            r = o.measure_device(ks, ws, profile=False)
            rr = reduce_workload(o, r, None, ks)
            per_config[name] = {"workload": o.wl["desc"], "value": rr["value"], "unit": "MP/s", "steps": ks, "warmup": ws,
                                "images_per_step_per_gpu": o.n, "encode_mpps": rr["encode_mpps"], "decode_mpps": rr["decode_mpps"],
                                "ms_per_step": rr["ms_per_step"], "bpp": rr["bpp"], "job": rr["job"],
                                "decode_ms_per_step_rank_min_max": rr["decode_ms_per_step_rank_min_max"]}
            o.close()
        except Exception as e:       # noqa: BLE001 -- a side pass never takes the headline down
            per_config[name] = {"error": f"{type(e).__name__}: {e}"}
            if world > 1:
                raise
    # the primary workload once more with weights TRAINED by the reference's own code (tests/golden/ckpt_A_trained.npz, llicti_A
    # shapes): real mixtures are peaky (spreads at the 0.11-level clamp), which is the harder case for the decoder's search
    trained = os.path.join(ROOT, "tests", "golden", "ckpt_A_trained.npz")
    if world == 1 and not args.no_per_config and not args.weights_npz and os.path.exists(trained) and WORKLOADS[primary]["cfg"] == "llicti_A.json":
        try:
            import copy
            a2 = copy.copy(args)
            a2.weights_npz = trained
            ks, ws = min(K, 3), min(Wm, 2)
            o = Workload(primary, a2, rank, world, local_rank, max_batches=ks + ws)
            rr = reduce_workload(o, o.measure_device(ks, ws, profile=False), None, ks)
            per_config[primary + "_trained_weights"] = {
                "workload": o.wl["desc"] + "; weights: tests/golden/ckpt_A_trained.npz (short-trained by the reference's mode: train)",
                "value": rr["value"], "unit": "MP/s", "steps": ks, "warmup": ws, "images_per_step_per_gpu": o.n,
                "encode_mpps": rr["encode_mpps"], "decode_mpps": rr["decode_mpps"], "ms_per_step": rr["ms_per_step"], "bpp": rr["bpp"]}
            o.close()
        except Exception as e:       # noqa: BLE001
            per_config[primary + "_trained_weights"] = {"error": f"{type(e).__name__}: {e}"}
    # the training step on the reference's training batch (configs/llicti_A.json of the reference: 32 patches
    # of 160 x 160): forward() + backward through the library, fp32; never fatal
    if world == 1 and not args.no_per_config:
        try:
            per_config["train_step"] = train_step_pass(local_rank, min(K, 5), min(Wm, 3))
        except Exception as e:       # noqa: BLE001
            per_config["train_step"] = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0 and not args.no_cpu:
            per_config["train_step"]["cpu_baseline"] = cpu_train_leg()     # a few seconds of host work
    if rank != 0:
        return

    cpu = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline_leg(primary)      # ~15-25 s of CPU work in a fresh process; never fatal

    line = {
        "metric": METRIC, "value": res["value"], "unit": "MP/s", "n_gpus": world, "steps": K, "warmup": Wm,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg_entry,
        "encode_mpps": res["encode_mpps"], "decode_mpps": res["decode_mpps"],
        "encode_ms_per_step": res["encode_ms_per_step"], "decode_ms_per_step": res["decode_ms_per_step"],
        "decode_ms_per_step_rank_min_max": res["decode_ms_per_step_rank_min_max"],
        "bpsp": res["bpsp"], "bpp": res["bpp"], "bpp_delta_vs_reference_streams": bpp_delta,
        "compressed_bytes_per_step": res["compressed_bytes_per_step"], "job": res["job"],
        "roofline": roof, "rooflines_all_kernels": allr, "kernel_ms_per_step": kernel_ms, "cnn_tflops": cnn_tflops,
        "internal_traffic_per_step": internal,
        "kernel_profile_pass": {"note": "per-kernel-class CUDA events (llicti_profile) over a second pass of the same K steps "
                                        "right after the timed region",
                                "encode_ms_per_step": p_enc / K, "decode_ms_per_step": p_dec / K},
        "cpu_baseline": cpu, "e2e": res["e2e"], "gpu_launches": res["gpu_launches"], "clocks": clk,
        "decode_stats_per_step": dstats, "per_config": per_config,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="", choices=[""] + sorted(WORKLOADS),
                    help="default: c2 on one GPU, c3 (sharded) on several")
    ap.add_argument("--images", type=int, default=0, help="override the job's images per GPU")
    ap.add_argument("--weights-npz", default="", help="state_dict as .npz instead of the synthetic stand-in weights (llicti_A shapes)")
    ap.add_argument("--batch", type=int, default=0, help="override the workload's images per batch (= per step)")
    ap.add_argument("--cnn", type=int, default=int(os.environ.get("LLICTI_CNN", "1")), help="0 fp32 CUDA cores, 1 tcgen05")
    ap.add_argument("--decode-impl", type=int, default=0, help="0 default schedules, 1 legacy one-warp-per-chain")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-train-leg", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-per-config", action="store_true", help="skip the short passes over the other configs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.cpu_train_leg:
        cpu_train_pass()
        return
    if args.impl == "reference":
        if not args.workload:
            args.workload = "c2" if world == 1 else "c3"
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU oracle)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
