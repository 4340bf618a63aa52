#!/usr/bin/env python
"""Differences of the mDBC correction to the reference dumps: how many corrected wall densities differ from the reference's, by how
many ulp, and what that does to Ace / Ar (the rounding variants this tool compared are recorded in profiles/r02_mdbc_rounding.log)."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import capi
import parity
from parity import Golden, rl2

lib = dsph.lib()
for variant in ["production"]:
    for name in parity.MDBC_CASES:
        g = Golden(name); cfg = g.cfg
        for step in g.steps:
            stages = [("fpre", capi.INTERSTEP_SYMPREDICTOR, "div0"), ("fcor", capi.INTERSTEP_SYMCORRECTOR, "div1")] if g.symplectic else [("fver", capi.INTERSTEP_VERLET, "div0")]
            for stg, inter, div in stages:
                ctx = parity.make_ctx(lib, cfg, 20000)
                s0 = g.stage(step, div)
                parity.upload_state(ctx, s0, cfg)
                ctx.celldiv(False)
                ref = g.stage(step, stg)
                ctx.forces(inter)
                ace, ar = ctx.forces_download()
                st = ctx.download()
                npb = int(ref["npb"]); npbok = int(ref["npbok"])
                a = st["velrhop"][:npb, 3]; b = ref["velrhop"][:npb, 3]
                nd = int(np.count_nonzero(a != b))
                ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))
                print("variant %s %-32s %3d %s  rhop: %d of %d differ (max %d ulp) rl2(dev) %.2e | ace %.2e ar %.2e"
                      % (variant, name, step, stg, nd, npbok, int(ulp.max()), rl2(a - float(cfg.rhopzero), b - float(cfg.rhopzero)), rl2(ace[npb:], ref["ace"][npb:]), rl2(ar, ref["ar"])))
                ctx.close()
