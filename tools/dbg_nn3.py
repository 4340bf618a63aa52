import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import capi
import parity
from parity import Golden
lib = dsph.lib()
g = Golden("nn3_2d_sym_fda_lam_ddt3_shift"); cfg = g.cfg
for step in g.steps:
    ctx = parity.make_ctx(lib, cfg, 20000)
    d1 = g.stage(step, "div1")
    parity.upload_state(ctx, d1, cfg)
    # pre arrays for the corrector: pospre / velrhoppre from the dump
    ctx.celldiv(False)
    ref = g.stage(step, "fcor")
    fr = ctx.forces(capi.INTERSTEP_SYMCORRECTOR)
    ace, ar = ctx.forces_download()
    print(step, "fcor ace", parity.rl2(ace, ref["ace"]), "ar", parity.rl2(ar, ref["ar"]))
    ctx.close()
    # chained like the test
    ctx = parity.make_ctx(lib, cfg, 20000)
    s0 = g.stage(step, "div0"); parity.upload_state(ctx, s0, cfg); ctx.celldiv(False)
    dt = float(s0["symplecticdtpre"])
    ctx.forces(capi.INTERSTEP_SYMPREDICTOR); ctx.step_sympre(dt); ctx.celldiv(True)
    ctx.forces(capi.INTERSTEP_SYMCORRECTOR); ctx.step_symcor(dt)
    u = g.stage(step, "ucor"); st = ctx.download()
    dpos = np.abs(st["pos"] - u["pos"]).max(axis=1)
    i = int(np.argmax(dpos))
    print(step, "max dpos", dpos[i], "particle", i, "idp", st["idp"][i], "code", hex(int(st["code"][i])), "pos", st["pos"][i], u["pos"][i], "n>1e-8:", int((dpos > 1e-8).sum()))
    sp = ref["shiftposfs"][i]; print("  ref shiftposfs", sp, "vel", u["velrhop"][i], st["velrhop"][i])
    up = g.stage(step, "div1")
    print("  pos before corr", up["pos"][i], "pospre", up["pospre"][i] if "pospre" in up else None, "dt", dt, "dp", cfg.dp, "maxdist", 0.1 * cfg.dp)
    ctx.close()
