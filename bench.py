#!/usr/bin/env python
"""Benchmark of the LLICTI compress/decompress hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4] [--impl b200|reference]

A step = one pass of the hot path (compress then decompress) over one batch of synthetic
images.  Default workload (N=1) is BASELINE.json configs[1]: llicti_B, 24 synthetic 768x512
images, torchac-compatible streams.  Rank 0 prints ONE JSON line on stdout.

  value     whole-job round-trip throughput (MP/s) with the batch already resident in HBM
  e2e       the same through llicti_encode_host / llicti_decode_host with pinned HOST buffers
            (host<->device copies inside the timed region)
  roofline  dominant kernel class: algorithmic bytes (or flops) / its CUDA-event time
  cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded sample

`--impl reference` times the CPU oracle alone (the reference is pure Python/PyTorch and cannot
travel to the GPU box; oracle/llicti_oracle.py keeps its cost structure).
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
    # name: (description, config, per-GPU images, H, W, sub_len)
    "c1": ("configs[1]: llicti_B eval_model, 24 synthetic 768x512 RGB images, torchac-compatible streams",
           "llicti_B.json", 24, 512, 768, 0),
    "c2": ("configs[2]: llicti_A, synthetic 2040x1356 images, interleaved-substream coder",
           "llicti_A.json", 25, 1356, 2040, 2048),
    "c3": ("configs[3]: llicti_A, synthetic 3840x2160 images, interleaved-substream coder",
           "llicti_A.json", 8, 2160, 3840, 2048),
    "c4": ("configs[4]: llicti_A, synthetic 512x512 images, interleaved-substream coder",
           "llicti_A.json", 256, 512, 512, 2048),
    "c0": ("configs[0]: llicti_A eval_model, 24 synthetic 768x512 RGB images, torchac-compatible streams",
           "llicti_A.json", 24, 512, 768, 0),
}
METRIC = "encode+decode round-trip megapixels/s (compress then decompres of every image)"
MAC_PER_POS = {88: (53152, 61600, 78496), 60: (29520, 35280, 46800)}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def synthetic_batch(n, H, W, seed0):
    """n distinct photographic-like images; a pool of 8 generated images is tiled with cheap
    per-image perturbations so that large batches do not take minutes of host time."""
    from llicti_b200.synth import synthetic_image
    pool = [synthetic_image(H, W, seed0 + i) for i in range(min(n, 8))]
    out = np.empty((n, 3, H, W), dtype=np.uint8)
    for i in range(n):
        img = pool[i % len(pool)]
        if i >= len(pool):
            img = np.roll(img, shift=(7 * i) % W, axis=2)
            img = np.clip(img.astype(np.int16) + ((i // len(pool)) % 5 - 2), 0, 255).astype(np.uint8)
        out[i] = img
    return out


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


def cpu_oracle_pass(cfg_json, H, W, steps, warmup, seed0=0):
    """Time the CPU oracle (compress + decompress of one image per step)."""
    from oracle import llicti_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    ocfg = O.OracleConfig.from_dict(cfg_json)
    codec = O.OracleCodec(ocfg, O.synthetic_state_dict(ocfg), sub_len=0)
    times = []
    for it in range(warmup + steps):
        img = O.synthetic_image(H, W, seed0 + it)
        t0 = time.perf_counter()
        bsl = codec.compress(img)
        t1 = time.perf_counter()
        rec = codec.decompress(bsl)
        t2 = time.perf_counter()
        assert np.array_equal(rec, img)
        log(f"[cpu oracle] step {it}: enc {t1 - t0:.2f}s dec {t2 - t1:.2f}s")
        if it >= warmup:
            times.append((t1 - t0, t2 - t1))
    enc = sum(t[0] for t in times)
    dec = sum(t[1] for t in times)
    px = H * W * len(times) / 1e6
    return {"value": px / (enc + dec), "encode_mpps": px / enc, "decode_mpps": px / dec,
            "ms_per_step": 1e3 * (enc + dec) / len(times), "cores": torch.get_num_threads()}


def run_reference(args, rank, world):
    if rank != 0:
        return
    desc, cfg_name, n_img, H, W, sub_len = WORKLOADS[args.workload]
    cfg = load_cfg(cfg_name)
    # the reference cannot hold a 4K image's dense CDF tables; time 768x512 crops for big shapes
    sh, sw = (H, W) if H * W <= 768 * 512 else (512, 768)
    r = cpu_oracle_pass(cfg, sh, sw, args.steps, args.warmup)
    sample = f"{args.steps} step(s) of 1 synthetic {sw}x{sh} image each, compress+decompres, {r['cores']} torch threads"
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "model_config": cfg_name, "sample": sample},
            "encode_mpps": r["encode_mpps"], "decode_mpps": r["decode_mpps"],
            "cpu_baseline": {"value": r["value"], "unit": "MP/s", "cores": r["cores"], "kind": "port", "sample": sample},
            "e2e": {"value": r["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# How each kernel class is bounded: the CNN by the tensor pipe, everything else is
# byte/integer work whose ceiling is HBM bandwidth -- the serial coder chains and the erfc-heavy
# CDF kernels sit far below it by nature (latency / instruction issue), which the fractions show.
BOUND = {"cnn": "tensor"}


def algorithmic_work(geom, chs, n, blob_bytes, decode_impl=0, piped=False):
    """Per-step algorithmic bytes / flops of each kernel class for n images."""
    S = geom.num_scales
    pos = [geom.Hs[s] * geom.Ws[s] for s in range(S)]
    sym_band = [[geom.crop_h[s][b] * geom.crop_w[s][b] for b in range(3)] for s in range(S)]
    macs = sum(pos[s] * sum(MAC_PER_POS[chs]) for s in range(S))
    coded_pos = sum(sum(sym_band[s]) for s in range(S))
    return {
        "split": n * (3 * geom.H * geom.W + 2 * 12 * sum(pos)),                 # u8 in, int16 planes out
        "cnn_flops": 2 * n * 2.0 * macs,                                        # the CNN runs in both directions
        "cnn": 2 * n * sum(pos[s] * (2 * 3 * (b + 1) + 240) for s in range(S) for b in range(3)),
        "bounds": n * coded_pos * (240 + 6 + 12),                               # 258 B per (position, band)
        "encode": n * geom.symbols * 4 + blob_bytes,                            # 4 B bounds in + bytes out
        "window": n * coded_pos * (240 + 6 + 3 * 64),                           # params + symbols in, 3 window rows out
        # split schedule: window rows + stream bytes in, symbols out; piped schedule: the one kernel also
        # produces the windows; legacy: params in, symbols out
        "decode": (n * coded_pos * ((240 + 6 + 3 * 64) * piped + 3 * 64 + 6) + blob_bytes) if decode_impl == 0
        else (n * coded_pos * (240 + 6) + blob_bytes),
        "merge": n * (2 * 12 * sum(pos) + 3 * geom.H * geom.W),
    }


def kernel_ms_has_no_window(prof):
    """True when the decode ran the piped schedule (windows produced inside the decode kernel)."""
    return prof.get("window", (0.0, 0))[1] == 0


def run_b200(args, rank, world, local_rank):
    from llicti_b200 import Codec, CodecConfig, _lib as L
    from llicti_b200 import synth           # synthetic weights / images; the oracle is used by the cpu_baseline leg only

    desc, cfg_name, n_img, H, W, sub_len = WORKLOADS[args.workload]
    if args.images:
        n_img = args.images
    cfg = load_cfg(cfg_name)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ccfg = CodecConfig.from_json_dict(cfg, sub_len=sub_len, numerics=L.NUM_TORCH_CUDA, cnn_impl=args.cnn,
                                      device=local_rank, decode_impl=args.decode_impl)
    sd = synth.synthetic_state_dict(ccfg.chs, ccfg.num_mixtures, int(cfg["Evens"][0]), int(cfg["Odds"][0]))
    codec = Codec(ccfg, sd)
    geom = codec.geometry(H, W)
    S = geom.num_scales
    st = 2 ** S
    from llicti_b200.shard import shard_range, reduce_stats
    # weak scaling: the job is world * n_img images, rank r codes its contiguous shard; no data-path collective
    first, last = shard_range(world * n_img, rank, world)
    assert last - first == n_img
    # Every shard holds the same synthetic image set: decode time depends on content (symbols outside their window take
    # the slow path; measured spread between image sets at N=8: 32 -> 38 ms per batch), and the weak-scaling number is
    # meant to show the system, not which rank drew the hardest pictures.
    rgb_h = torch.from_numpy(synthetic_batch(n_img, H, W, 1000)).pin_memory()
    rgb_np = rgb_h.numpy()
    rgb_d = rgb_h.to(dev)
    x00_np = np.ascontiguousarray(rgb_np[:, :, ::st, ::st])
    x00_d = torch.from_numpy(x00_np).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    out_pinned = torch.empty(n_img * int(geom.max_stream_bytes), dtype=torch.uint8).pin_memory().numpy()
    rec_pinned = torch.empty((n_img, 3, H, W), dtype=torch.uint8).pin_memory().numpy()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    enc_out = [None]                      # device output buffers, allocated by the first (untimed) step
    rec_out = [None]

    def dev_step(timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        flush.zero_()
        ev[0].record()
        enc_out[0] = codec.encode_dev(rgb_d, enc_out[0])
        blob, off, mm = enc_out[0]
        ev[1].record()
        flush.zero_()                      # decode starts cold as well (outside both timed spans)
        ev2 = torch.cuda.Event(enable_timing=True)
        ev2.record()
        if ENCODE_ONLY:                    # kernel experiments whose streams are not decodable (tools/)
            ev[2].record()
            torch.cuda.synchronize()
            return ev[0].elapsed_time(ev[1]), 1e-3, blob, off, rgb_d
        rec_out[0] = rec = codec.decode_dev(blob, off, mm, x00_d, n_img, H, W, rec_out[0])
        ev[2].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[1]), ev2.elapsed_time(ev[2]), blob, off, rec

    def host_step():
        if ENCODE_ONLY:
            return 1e-3, 1e-3, 0, 0, None, rgb_np
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
        e0.record()
        blob, off, mm = codec.encode_host(rgb_np, out_pinned)
        e1.record()
        flush.zero_()
        e2.record()
        rec = codec.decode_host(blob, off, mm, x00_np, n_img, H, W, rec_pinned)
        e3.record()
        torch.cuda.synchronize()
        h2d = rgb_np.nbytes + blob.nbytes + off.nbytes + mm.nbytes + x00_np.nbytes
        d2h = blob.nbytes + off.nbytes + mm.nbytes + rec.nbytes
        return e0.elapsed_time(e1), e2.elapsed_time(e3), h2d, d2h, blob, rec

    # ---- warm-up ---------------------------------------------------------------------------
    for _ in range(args.warmup):
        dev_step(False)
    enc_ms, dec_ms, blob, off, rec = dev_step(False)
    assert torch.equal(rec, rgb_d), "device round trip is not lossless"
    blob_bytes = int(off[-1].item())

    # ---- timed region: device-resident -------------------------------------------------------
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    launches0 = codec.launches
    codec.decode_stats()
    t_enc = t_dec = 0.0
    for _ in range(args.steps):
        a, b, *_ = dev_step(True)
        t_enc += a
        t_dec += b
    barrier()
    dstats = codec.decode_stats()
    launches = codec.launches - launches0

    # ---- the same K steps once more with the library's per-kernel-class events (they need eager launches: the decode
    #      of the timed region above is replayed as one CUDA graph) ----------------------------------------------
    codec.profile(True)
    p_enc = p_dec = 0.0
    for _ in range(args.steps):
        a, b, *_ = dev_step(True)
        p_enc += a
        p_dec += b
    prof = codec.profile_read()
    codec.profile(False)
    codec.decode_stats()
    barrier()

    # ---- rate of the substream container against the torchac-compatible streams (= the reference's bytes) of the
    #      same images, untimed, on a few images of the batch ----------------------------------------------------
    bpp_delta = None
    if sub_len > 0 and rank == 0 and not ENCODE_ONLY:
        m = min(n_img, 4)
        ns = 9 * S
        compat = Codec(CodecConfig.from_json_dict(cfg, sub_len=0, numerics=L.NUM_TORCH_CUDA, cnn_impl=args.cnn,
                                                  device=local_rank, decode_impl=args.decode_impl), sd)
        _, c_off, _ = compat.encode_dev(rgb_d[:m].contiguous())
        compat_bytes = int(c_off[-1].item())
        sub_bytes = int(off[m * ns].item()) + 5 * m          # + the container's mode tag in the header row
        compat.close()
        del compat
        torch.cuda.empty_cache()
        bpp_delta = {"value": (sub_bytes - compat_bytes) / compat_bytes, "images": m, "substream_bytes": sub_bytes,
                     "torchac_compatible_bytes": compat_bytes,
                     "note": "(substream container - torchac-compatible streams) / torchac-compatible streams, all "
                             "length tables and flush bytes counted; the compatible streams are the reference's bytes"}

    # ---- timed region: end to end through host buffers ------------------------------------------
    for _ in range(max(1, min(args.warmup, 2))):
        host_step()
    barrier()
    h_enc = h_dec = 0.0
    for _ in range(args.steps):
        a, b, h2d, d2h, hblob, hrec = host_step()
        h_enc += a
        h_dec += b
    barrier()
    clk = clocks.stop()        # sampled over both timed regions (device-resident and end-to-end)
    assert np.array_equal(hrec, rgb_np), "host round trip is not lossless"

    # ---- reduce over ranks (max time, summed work) ---------------------------------------------
    (px_total, bytes_total, launches_total, h2d_total, d2h_total), (t_enc, t_dec, h_enc, h_dec) = reduce_stats(
        [n_img * H * W, blob_bytes, launches, h2d, d2h], [t_enc, t_dec, h_enc, h_dec], device=dev)   # NCCL: rate statistics only
    if rank != 0:
        return
    K = args.steps
    mp = px_total / 1e6
    value = mp * K / ((t_enc + t_dec) / 1e3)
    e2e = mp * K / ((h_enc + h_dec) / 1e3)

    # ---- roofline of the dominant kernel class (rank 0's events) ------------------------------------
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        pk = json.load(open(pk_path))
        peaks = {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops_sustained": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]),
                 "src": "measured"}
    piped = sub_len == 0 and kernel_ms_has_no_window(prof)
    work_step = algorithmic_work(geom, ccfg.chs, n_img, blob_bytes, args.decode_impl, piped)
    kernel_ms = {k: v[0] / K for k, v in prof.items()}
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        traffic_db = json.load(open(tpath)).get(args.workload, {})
    coded_symbols = n_img * geom.symbols

    def roofline_of(cls):
        """Algorithmic work of one average launch group / its average duration (CUDA events on the
        launching stream during the timed region)."""
        groups = max(prof[cls][1] / K, 1.0)
        ms = kernel_ms[cls] / groups
        if ms <= 0:
            return None
        if BOUND.get(cls) == "tensor" and args.cnn == 1:
            ach = work_step["cnn_flops"] / groups / (ms * 1e-3) / 1e12
            r = {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s"}
        else:
            ach = work_step[cls] / groups / (ms * 1e-3) / 1e9
            r = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
        r["frac"] = r["achieved"] / r["peak"]
        t = traffic_db.get(cls)      # dram bytes of one captured launch (ncu --set full), scaled by work units to the average launch
        r["traffic"] = (t["dram_bytes"] / t["units"] * (n_img * t["units_per_image_step"] / groups)) if t else None
        r.update({"kernel": cls, "launch_groups_per_step": groups, "ms_per_launch_group": ms,
                  "share_of_step": kernel_ms[cls] / ((p_enc + p_dec) / K)})
        if t:
            r["traffic_source"] = t["source"]
        return r

    dom = max(kernel_ms, key=kernel_ms.get)
    roof = roofline_of(dom)
    roof["peak_source"] = peaks["src"]
    rooflines = {c: {k: v for k, v in r.items() if k in ("bound", "achieved", "peak", "unit", "frac", "share_of_step")}
                 for c in kernel_ms if kernel_ms[c] > 0 and c in work_step for r in [roofline_of(c)] if r}
    cnn_tflops = work_step["cnn_flops"] / (max(kernel_ms["cnn"], 1e-9) * 1e-3) / 1e12

    # ---- CPU baseline: the oracle on a bounded sample (rank 0, N=1 only) -----------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        sh, sw = (H, W) if H * W <= 768 * 512 else (512, 768)
        r = cpu_oracle_pass(cfg, sh, sw, 4, 0)      # ~15-20 s of CPU work
        cpu = {"value": r["value"], "unit": "MP/s", "cores": r["cores"], "kind": "port",
               "sample": f"4 synthetic {sw}x{sh} images, compress+decompres of each once, {r['cores']} torch threads",
               "encode_mpps": r["encode_mpps"], "decode_mpps": r["decode_mpps"]}

    line = {
        "metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": (t_enc + t_dec) / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "model_config": cfg_name, "images_per_gpu": n_img, "height": H, "width": W,
                   "sub_len": sub_len, "cnn_impl": "tcgen05" if args.cnn == 1 else "fp32-cuda-core",
                   "decode_impl": "windows+chains" if args.decode_impl == 0 else "legacy-warp",
                   "weights": "llicti_b200.synth.synthetic_state_dict(seed=1337) (shipped checkpoint absent)",
                   "l2": "256 MiB buffer written before every timed encode and decode (L2 flushed)",
                   "parallelism": f"images sharded over {world} GPU(s), no data-path collective; every shard is the same "
                                  f"{n_img}-image synthetic set (identical work per GPU)"},
        "encode_mpps": mp * K / (t_enc / 1e3), "decode_mpps": mp * K / (t_dec / 1e3),
        "encode_ms_per_step": t_enc / K, "decode_ms_per_step": t_dec / K,
        "bpsp": bytes_total * 8 / (px_total * 3), "bpp": bytes_total * 8 / px_total, "bpp_delta_vs_reference_streams": bpp_delta,
        "compressed_bytes_per_step": bytes_total,
        "roofline": roof, "rooflines_all_kernels": rooflines, "kernel_ms_per_step": kernel_ms, "cnn_tflops": cnn_tflops,
        "kernel_profile_pass": {"note": "per-kernel-class CUDA events (llicti_profile) over a second pass of the same K steps "
                                        "right after the timed region; the timed region replays the decode as one CUDA graph",
                                "encode_ms_per_step": p_enc / K, "decode_ms_per_step": p_dec / K},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e, "unit": "MP/s", "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": d2h_total,
                "encode_mpps": mp * K / (h_enc / 1e3), "decode_mpps": mp * K / (h_dec / 1e3),
                "api": "llicti_encode_host + llicti_decode_host, pinned host buffers"},
        "gpu_launches": int(launches_total), "clocks": clk,
        "decode_stats_per_step": {k: v / K for k, v in dstats.items()},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=0, help="override images per GPU")
    ap.add_argument("--cnn", type=int, default=int(os.environ.get("LLICTI_CNN", "1")), help="0 fp32 CUDA cores, 1 tcgen05")
    ap.add_argument("--decode-impl", type=int, default=0, help="0 windows + serial chains, 1 legacy one-warp-per-chain")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
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
