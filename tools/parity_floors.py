#!/usr/bin/env python
"""Prints, for every golden dump and force stage, the raw differences of the CUDA path to the reference next to the fp32 conditioning
floors of parity.py (tool for deriving / re-checking the constants stated in DESIGN.md "Parity tolerances")."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import capi
import parity
from parity import Golden


def noise_floors(cfg, nfluid, nall):
    """fp32 conditioning floors (L2 over the compared particles) below which two correct implementations cannot agree:
    ace: Tait's equation P = B((rho/rho0)^gamma - 1) is evaluated where the power is ~1, so a rounding error of u ulp of the power
         (2^-23) is an absolute pressure error u*2^-23*B per particle, i.e. an acceleration error ~ u*2^-23*B/(rho0*h) (pressure
         gradient over one smoothing length); u = 0.29 = rms rounding error of ONE correctly rounded result (uniform in +-1/2 ulp).
    ar:  the density-diffusion term sums delta*h*cs0*(rho_j-rho_i)... over the neighbours: one ulp of rho (rho0*2^-23) in every
         difference is a rate error ~ delta*cs0*rho0*2^-23/h per particle (same u).
    Multiphase cases use the constants of the reference phase (cfg.cteb / cfg.rhopzero)."""
    ulp = 2.0 ** -23
    u = 0.29
    h = float(cfg.kernelh); rho0 = float(cfg.rhopzero); b = float(cfg.cteb); gamma = float(cfg.gamma)
    cs0 = (b * gamma / rho0) ** 0.5
    f_ace = u * (nfluid ** 0.5) * ulp * b / (rho0 * h)
    f_ar = u * (nall ** 0.5) * max(float(cfg.ddtvalue), 0.1) * cs0 * rho0 * ulp / h
    return f_ace, f_ar



lib = dsph.lib()
for name in parity.GOLDEN_CASES:
    g = Golden(name)
    cfg = g.cfg
    for step in g.steps:
        stages = [("fpre", capi.INTERSTEP_SYMPREDICTOR, "div0"), ("fcor", capi.INTERSTEP_SYMCORRECTOR, "div1")] if g.symplectic else [("fver", capi.INTERSTEP_VERLET, "div0")]
        for stg, inter, div in stages:
            if not g.has(step, stg) or not g.has(step, div):
                continue
            ctx = parity.make_ctx(lib, cfg, 20000)
            s0 = g.stage(step, div)
            parity.upload_state(ctx, s0, cfg)
            ctx.celldiv(False)
            ref = g.stage(step, stg)
            ctx.forces(inter)
            ace, ar = ctx.forces_download()
            npb = int(ref["npb"])
            if ace.shape[0] != ref["ace"].shape[0]:
                ctx.close(); continue
            da = np.linalg.norm((ace[npb:] - ref["ace"][npb:]).astype(np.float64)); na = np.linalg.norm(ref["ace"][npb:].astype(np.float64))
            dr = np.linalg.norm((ar - ref["ar"]).astype(np.float64)); nr = np.linalg.norm(ref["ar"].astype(np.float64))
            fa, fr = noise_floors(cfg, ace.shape[0] - npb, ar.shape[0])
            print("%-30s %3d %s  ace: d %.3e ref %.3e rel %.2e floor %.3e d/floor %.2f | ar: d %.3e ref %.3e rel %.2e floor %.3e d/floor %.2f"
                  % (name, step, stg, da, na, da / na, fa, da / fa, dr, nr, dr / max(nr, 1e-30), fr, dr / fr))
            ctx.close()
