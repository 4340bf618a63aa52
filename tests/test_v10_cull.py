"""Conservativeness of the half-precision candidate culling of the production force kernel (csrc/forces_v10.cuh), on the CPU.

The kernel tests candidates with three HFMA2 on a packed-half copy of the cell-relative positions and accepts when
n - 2 o.p <= (2h)^2 + V10_MARGIN - |o|^2 (scaled units: cell = [-1,1]^3).  The pair body then applies the exact fp32 test, so a
FALSE POSITIVE only costs a wasted slot, but a FALSE NEGATIVE would drop a real pair: this test emulates the half pipeline
bit-exactly in numpy (fp16 FMA = exact product and sum in float64, one rounding to float16) on millions of pairs placed at and
just inside the cut-off, in every neighbour-cell configuration, and checks that no real pair is rejected -- with the margin the
kernel uses and with HALF of it (safety factor 2).  It also reports the false-positive rate the margin costs."""
import re
import os

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def kernel_margin():
    src = open(os.path.join(ROOT, "dualsphysics_multilayer_b200", "csrc", "forces_v10.cuh")).read()
    return float(re.search(r"V10_MARGIN=([0-9.eE+-]+)f", src).group(1))


def f16(x):
    return np.asarray(x, np.float64).astype(np.float16)


def fma16(a, b, c):
    """a*b+c on float16 inputs with ONE rounding (HFMA2): float64 holds the exact product and sum of these magnitudes."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float16)


def f16_round_up(x32):
    h = x32.astype(np.float16)
    low = h.astype(np.float32) < x32
    return np.where(low, np.nextafter(h, np.float16(np.inf)), h).astype(np.float16)


def cull(pp1, d, pp2, scell, ks2, margin):
    """The device computation: pp1/pp2 float32 cell-relative positions, d = neighbour-cell offset (-1,0,1)^3 of pp2's cell."""
    hs = np.float32(2.0) / np.float32(scell)
    # candidate side (celldiv.cuh, ker_div_reorder)
    h2 = f16(pp2 * hs - np.float32(1.0))
    f2 = h2.astype(np.float32)
    n2 = f16(f2[:, 0] * f2[:, 0] + f2[:, 1] * f2[:, 1] + f2[:, 2] * f2[:, 2])
    # lane side (forces_v10.cuh)
    b = (pp1 * hs - np.float32(1.0)).astype(np.float32)
    os_ = (b - np.float32(2.0) * d.astype(np.float32)).astype(np.float32)
    m = f16(np.float32(-2.0) * os_)
    thr0 = np.float32(ks2) * hs * hs + np.float32(margin)
    thr = f16_round_up(((thr0 - os_[:, 2] * os_[:, 2] - os_[:, 1] * os_[:, 1]) - os_[:, 0] * os_[:, 0]).astype(np.float32))
    s = fma16(m[:, 0], h2[:, 0], n2)
    s = fma16(m[:, 1], h2[:, 1], s)
    s = fma16(m[:, 2], h2[:, 2], s)
    return s <= thr


def make_pairs(rng, n, scell, rlo, rhi):
    """p2 uniform in its cell, p1 at distance r in [rlo,rhi]*(scell) from it (any direction); returns cell-relative floats."""
    p2 = rng.random((n, 3)) * scell
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    r = (rlo + (rhi - rlo) * rng.random(n)) * scell
    p1 = p2 + u * r[:, None]                       # in the frame of p2's cell: p1 lies in [-scell, 2*scell)
    c1 = np.floor(p1 / scell)                      # cell of p1 relative to p2's cell
    pp1 = (p1 - c1 * scell).astype(np.float32)
    pp2 = p2.astype(np.float32)
    d = (-c1).astype(np.int32)                     # offset of p2's cell as seen from p1's cell
    ok = np.all(np.abs(d) <= 1, axis=1) & np.all(pp1 < np.float32(scell), axis=1)
    true_r2 = ((pp1.astype(np.float64) - d * scell - pp2.astype(np.float64)) ** 2).sum(axis=1)
    return pp1[ok], d[ok], pp2[ok], true_r2[ok]


def test_half_culling_never_drops_a_real_pair():
    margin = kernel_margin()
    rng = np.random.default_rng(2024)
    worst = -1e9
    for scell in (0.0069282, 0.0294449, 1.0, 0.00173205):          # 2h of several resolutions (coefh=1: 2*sqrt(3)*dp)
        ks2 = np.float32(scell) * np.float32(scell)                 # cell = 2h -> (2h)^2 = scell^2
        for rlo, rhi in ((0.98, 1.0), (0.999, 1.0), (0.0, 1.0)):
            pp1, d, pp2, r2 = make_pairs(rng, 400_000, scell, rlo, rhi)
            real = r2 <= float(ks2) * (1 + 1e-6)
            for mg in (margin, 0.5 * margin):
                hit = cull(pp1, d, pp2, scell, ks2, mg)
                assert np.all(hit[real]), ("real pair rejected", scell, rlo, mg, int(np.count_nonzero(real & ~hit)))
            # how much margin the worst real pair actually needed (scaled units)
            need = cull_slack(pp1[real], d[real], pp2[real], scell, ks2)
            worst = max(worst, float(need.max()))
    print("largest margin needed by a real pair: %.4f (kernel uses %.4f)" % (worst, margin))
    assert worst < 0.5 * margin


def cull_slack(pp1, d, pp2, scell, ks2):
    """s - thr(margin=0): the margin a pair needs to pass."""
    hs = np.float32(2.0) / np.float32(scell)
    h2 = f16(pp2 * hs - np.float32(1.0)); f2 = h2.astype(np.float32)
    n2 = f16(f2[:, 0] * f2[:, 0] + f2[:, 1] * f2[:, 1] + f2[:, 2] * f2[:, 2])
    b = (pp1 * hs - np.float32(1.0)).astype(np.float32)
    os_ = (b - np.float32(2.0) * d.astype(np.float32)).astype(np.float32)
    m = f16(np.float32(-2.0) * os_)
    thr = ((np.float32(ks2) * hs * hs - os_[:, 2] * os_[:, 2] - os_[:, 1] * os_[:, 1]) - os_[:, 0] * os_[:, 0]).astype(np.float64)
    s = fma16(m[:, 0], h2[:, 0], n2); s = fma16(m[:, 1], h2[:, 1], s); s = fma16(m[:, 2], h2[:, 2], s)
    return s.astype(np.float64) - thr


def test_half_culling_false_positive_rate_is_small():
    """Candidates uniformly distributed around a particle (the 27-cell neighbourhood): the share of accepted candidates that the
    exact test then rejects must stay at the per-cent level (it costs pair-round slots, not correctness)."""
    margin = kernel_margin()
    rng = np.random.default_rng(7)
    scell = 0.0069282
    n = 1_500_000
    pp1 = np.repeat((rng.random((1, 3)) * scell).astype(np.float32), n, axis=0)
    d = rng.integers(-1, 2, size=(n, 3)).astype(np.int32)
    pp2 = (rng.random((n, 3)) * scell).astype(np.float32)
    r2 = ((pp1.astype(np.float64) - d * scell - pp2.astype(np.float64)) ** 2).sum(axis=1)
    ks2 = np.float32(scell) * np.float32(scell)
    hit = cull(pp1, d, pp2, scell, ks2, margin)
    real = r2 <= float(ks2)
    assert np.all(hit[real])
    fp = np.count_nonzero(hit & ~real) / max(1, np.count_nonzero(real))
    print("accepted/real = %.4f (false positives %.2f %% of the real pairs); real share of candidates %.3f" % (hit.sum() / real.sum(), 100 * fp, real.mean()))
    assert fp < 0.05
