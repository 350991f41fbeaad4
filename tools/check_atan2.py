#!/usr/bin/env python3
"""Error bound of atan2_q (csrc/velo_common.cuh) emulated in float32 with the same operation order, against f64 atan2."""
import numpy as np
f = np.float32
c = [f(0.9999772191230201), f(-0.3326228279880411), f(0.1935403746008874), f(-0.11642647568056484), f(0.05264734316289409), f(-0.011719132173485577)]

def atan2_q(y, x):
    ax, ay = np.abs(x), np.abs(y)
    mn, mx = np.minimum(ax, ay), np.maximum(ax, ay)
    a = np.where(mx > 0, mn / np.where(mx > 0, mx, 1), 0).astype(f)
    s = (a * a).astype(f)
    p = np.full_like(a, c[5])
    for k in range(4, -1, -1): p = (p.astype(np.float64) * s.astype(np.float64) + np.float64(c[k])).astype(f)      # fmaf: the product is exact in f64
    r = (p * a).astype(f)
    r = np.where(ay > ax, (f(np.pi / 2) - r).astype(f), r)
    r = np.where(x < 0, (f(np.pi) - r).astype(f), r)
    return np.where(y < 0, -r, r).astype(f)

def max_error(n=1_000_000, seed=1):
    rng = np.random.default_rng(seed); worst = 0.0
    for scale in (1e-3, 1.0, 100.0, 1e4):
        x = (rng.standard_normal(n) * scale).astype(f)
        y = (rng.standard_normal(n) * scale * rng.choice([1e-3, 1, 1e3], n)).astype(f)
        worst = max(worst, np.abs(atan2_q(y, x).astype(np.float64) - np.arctan2(y.astype(np.float64), x.astype(np.float64))).max())
    th = np.linspace(-np.pi, np.pi, 2 * n + 1); x = (np.cos(th) * 17.3).astype(f); y = (np.sin(th) * 17.3).astype(f)
    return max(worst, np.abs(atan2_q(y, x).astype(np.float64) - np.arctan2(y.astype(np.float64), x.astype(np.float64))).max())

if __name__ == "__main__":
    e = max_error(4_000_000); print("max |atan2_q - atan2| =", e); assert e < 2.5e-6
