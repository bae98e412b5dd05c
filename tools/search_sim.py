#!/usr/bin/env python
"""CPU model of the lane decoder's symbol search (kernels_decode.cu: the Newton / Halley iteration on the stand-in CDF) on
mixtures the TRAINED network produces: how many stand-in evaluations a lane needs and -- what sets the kernel's time -- the
maximum over the 30 lanes of a warp.  A development aid for the search heuristics; not product code, not the oracle.

    python tools/search_sim.py [--variant base|...] [--n 30000]
"""
import argparse
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
import numpy as np                      # noqa: E402
from scipy.special import erfc          # noqa: E402


def mixtures(trained=True, n=30000, seed=0):
    """(sigma, mu, w) [n, 5] of the Y channel of band 0, scale 0, and the true symbol-space size, from a synthetic photo."""
    from oracle import llicti_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import trained_state_dict
    ocfg = O.OracleConfig()
    sd = trained_state_dict() if trained else O.synthetic_state_dict(ocfg)
    img = O.synthetic_image(256, 384, 321, noise=0.7)
    codec = O.OracleCodec(ocfg, sd)
    dump = O.StageDump()
    codec.compress(img, dump)
    prm = dump.params[(0, 0)]                     # [60, Hs, Ws]
    M = 5
    sig, mu, w = prm[0:M], prm[3 * M:4 * M], prm[6 * M:7 * M]            # Y channel
    rng = np.random.default_rng(seed)
    idx = rng.choice(sig[0].size, size=min(n, sig[0].size), replace=False)
    f = lambda a: a.reshape(M, -1)[:, idx].T.astype(np.float64)
    sig, mu, w = np.maximum(f(sig), 0.11 / 255), f(mu), np.maximum(f(w), 1e-6)
    w = w / (w.sum(1, keepdims=True) + 1e-9)
    return sig, mu, w


def q_of(sig, mu, w, x, scale, min_val):
    p = (min_val + x - 0.5) / 255.0
    z = (p - mu) / sig
    c = 0.5 * erfc(-z / np.sqrt(2))
    pdf = np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
    q = x + scale * (w * c).sum()
    dq = 1.0 + scale / 255.0 * (w * pdf / sig).sum()
    ddq = scale / 255.0 ** 2 * (w * pdf / sig / sig * (-z)).sum()
    return q, dq, ddq


def search(sig, mu, w, tf, variant, last=256, min_val=-127, margin=0.2):
    scale = 65536.0 - last
    mean = (w * mu).sum()
    sd = np.sqrt(max((w * (sig * sig + mu * mu)).sum() - mean * mean, 1e-12))
    smooth = np.where(w >= 0.004, sig, 1.0).min() >= 0.75 / 255
    u = min(max(tf / 65536.0, 1e-6), 1 - 1e-6)
    y = (np.log2(u) - np.log2(1 - u)) * 0.43436
    z0 = y / (1 + 0.044715 * y * y)
    zq = z0 - (z0 + 0.044715 * z0 ** 3 - y) / (1 + 0.134145 * z0 * z0)
    k = int(np.floor((mean + zq * sd) * 255 + 0.5)) - min_val
    d = int(np.argmax(w))
    if (variant.startswith("dom") or "first" in variant) and not smooth:
        k = int(np.floor((mu[d] + zq * sig[d]) * 255 + 0.5)) - min_val
    a, b, its = 0, last, 0
    for it in range(64):
        if b - a <= 1:
            return a, its
        k = min(max(k, a + 1), b - 1)
        q, dq, ddq = q_of(sig, mu, w, k, scale, min_val)
        its += 1
        res = tf - q
        dxn = res / dq
        dx = res / max(dq + 0.5 * ddq * dxn, 0.3 * dq)
        if res >= 0:
            a = k
            if b == k + 1 or (smooth and dx < 1 - margin):
                return k, its
        else:
            b = k
            if a == k - 1 or (smooth and dx >= -1 + margin):
                return k - 1, its
        kn = k + int(np.floor(min(max(dx + margin if res >= 0 else dx, -70000), 70000)))
        if variant in ("base", "baseclamp", "baseclamplin") and (not smooth) and abs(dx) > 3:
            best = b if res >= 0 else a
            for m in range(5):
                km = int(np.floor(mu[m] * 255 + 0.5)) - min_val + (0 if res >= 0 else 1)
                if res >= 0:
                    if k < km < best:
                        best = km
                elif best < km < k:
                    best = km
            if a < best < b:
                kn = best
        if variant.startswith("dom") and not smooth:
            from scipy.special import ndtr, ndtri
            p_k = (min_val + k - 0.5) / 255.0
            cd = ndtr((p_k - mu[d]) / sig[d])
            tgt = cd + res / (scale * w[d])
            if 1e-4 < tgt < 1 - 1e-4:
                xn = (mu[d] + sig[d] * ndtri(tgt)) * 255 + 0.5 - min_val      # index whose sampling point is the root
                kn = int(np.floor(xn))
                if variant == "dom2" and kn == k and res >= 0:
                    kn = k + 1
                if variant == "dom2" and kn == k and res < 0:
                    kn = k - 1
        if "clamp" in variant and not smooth and not ("dq" in variant and dq <= float(variant.split("dq")[1] or 4)):
            lim = 2 * (it + 1) if "lin" in variant else (1 << it) if "one" in variant else (3 << it) if "three" in variant else 2 << it
            kn = min(max(kn, k - lim), k + lim)
            if kn == k:
                kn = k + 1 if res >= 0 else k - 1
        k = (a + b) >> 1 if (kn <= a or kn >= b or it >= 8) else kn
    return a, its


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variant", default="base")
    ap.add_argument("--n", type=int, default=6000)
    ap.add_argument("--synthetic", action="store_true")
    args = ap.parse_args()
    sig, mu, w = mixtures(trained=not args.synthetic, n=args.n)
    rng = np.random.default_rng(1)
    its = np.zeros(len(sig), int)
    wrong = 0
    for i in range(len(sig)):
        tf = rng.uniform(0, 65536)
        g, its[i] = search(sig[i], mu[i], w[i], tf, args.variant)
    print("sigma (levels) quantiles of the narrowest weighted component:",
          np.round(np.quantile(np.where(w >= 0.004, sig, 1.0).min(1) * 255, [0.05, 0.25, 0.5, 0.75, 0.95]), 3))
    print("evaluations per symbol: mean %.2f, histogram %s" % (its.mean(), np.bincount(its)[:14]))
    groups = its[:len(its) // 30 * 30].reshape(-1, 30)
    print("max over 30 lanes: mean %.2f, histogram %s" % (groups.max(1).mean(), np.bincount(groups.max(1))[:14]))


if __name__ == "__main__":
    main()
