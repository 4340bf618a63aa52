#!/usr/bin/env python
"""Times the force-kernel variants on a developed dam-break state (device events).  usage: forcebench.py NP PRESTEPS [variants] [kind] [key=value ...]   (kind: dambreak | nn2layer; e.g. velgrad=2 tvisco=3)"""
import ctypes as C, sys, os, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import dualsphysics_multilayer_b200 as dsph
npart = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 20
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2]
L = dsph.lib(); l = L.lib
kind = sys.argv[4] if len(sys.argv) > 4 else "dambreak"
cd = dsph.make_case(kind, npart)
for kv in sys.argv[5:]:
    k, v = kv.split("=")
    k = {"velgrad": "nn_velgrad", "ddt": "tdensity", "shifting": "shiftmode"}.get(k, k)
    setattr(cd.cfg, k, type(getattr(cd.cfg, k))(float(v)))
ctx = dsph.create_context(cd)
ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, cd.codes())
ctx.celldiv(True)
for _ in range(pre):
    ctx.step()
fr, fc, br, bc = ctx.pips()
npf = ctx.np_ - ctx.npb
print("np=%d npf=%d pairs/fluid: checked %.1f real %.1f" % (ctx.np_, npf, fc / npf, fr / npf))
ref = None
for v in variants:
    l.dsph_set_force_variant(ctx.h, v)
    ts = []
    for it in range(4):
        ctx.celldiv(False)
        r = ctx.forces()
        ts.append(l.dsph_last_force_ms(ctx.h))
    ace, ar = ctx.forces_download()
    if ref is None:
        ref = (ace, ar)
    e1 = np.linalg.norm((ace - ref[0]).astype(np.float64)) / np.linalg.norm(ref[0].astype(np.float64))
    e2 = np.linalg.norm((ar - ref[1]).astype(np.float64)) / np.linalg.norm(ref[1].astype(np.float64))
    t = min(ts[1:])
    gb = (58.0 * npf + 16.0 * (fc + fr)) / (t * 1e-3) / 1e9
    print("variant %d: forces %.3f ms (all %s)  %.1f M fluid-evals/s  gather-model %.0f GB/s  vs v%d: ace %.2e ar %.2e dt %.9g"
          % (v, t, ["%.2f" % x for x in ts], npf / t / 1e3, gb, variants[0], e1, e2, r.dt))
