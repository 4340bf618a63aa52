"""Moving boundaries, host side (no GPU): the motion laws the library evaluates for the stand-alone driver (csrc/host/motion.cpp) against what
the reference's JDsMotion / JMotion computed at EVERY step of its own runs (the per-step log of the harness in tests/golden/mv*.npz)."""
import os
import tempfile

import numpy as np
import pytest

import parity
from parity import Golden

import dualsphysics_multilayer_b200 as dsph


@pytest.mark.parametrize("name", parity.MOVING_CASES + parity.MOVING_MDBC_CASES)
def test_motion_laws_match_reference_log(name):
    g = Golden(name)
    tdt, val = parity.motion_log(g)
    assert tdt.shape[0] == g.nsteps
    with tempfile.TemporaryDirectory() as tmp:
        parity.write_case_xml(g, tmp)
        motion = dsph.CaseMotion(tmp, "Case")
    assert motion.nobj == val.shape[1] == 1
    nmoved = 0
    worst = 0.0
    for s in range(g.nsteps):
        objs, active = motion.process(float(tdt[s, 0]), float(tdt[s, 1]))
        for i, d in enumerate(objs):
            ref = val[s, i]
            assert d.type == int(ref[0]), (name, s, d.type, ref[0])
            assert d.idbegin == int(g.cfg.casenfixed) and d.count == int(g.cfg.casenmoving)
            if d.type == 1:
                got = np.array(list(d.linmov) + list(d.linvel)); want = ref[1:7]
            elif d.type == 2:
                got = np.array(list(d.matmov)); want = ref[7:23]
            else:
                continue
            nmoved += 1
            # same formulas in double: differences are rounding of the last bits of sin / cos and of the matrix products
            err = float(np.abs(got - want).max() / max(1e-300, np.abs(want).max()))
            worst = max(worst, err)
            assert err < 1e-12, (name, s, got, want)
        assert active == any(int(val[s, i, 0]) != 0 for i in range(val.shape[1]))
    print(name, "steps with motion:", nmoved, "worst relative difference:", worst)
    assert nmoved >= 5
    if name == "mv_sym_gate_rect_wait":          # the gate stops after 0.004 s: the chain mvrect -> wait ends inside a step
        assert nmoved < g.nsteps
    motion.close()


def test_motion_xml_refuses_unsupported():
    xml = """<case><execution><particles np="4" nb="2" nbf="1" mkboundfirst="11" mkfluidfirst="1">
      <fixed mkbound="0" mk="11" begin="0" count="1" /><moving mkbound="1" mk="12" begin="1" count="1" refmotion="0" />
      <fluid mkfluid="0" mk="1" begin="2" count="2" /></particles>
      <motion><objreal ref="0"><begin mov="1" start="0" /><mvspline id="1" duration="1"><vel x="1" y="0" z="0" /></mvspline></objreal></motion>
      </execution></case>"""
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "Case.xml"), "w").write(xml)
        with pytest.raises(dsph.SphError):
            dsph.CaseMotion(tmp, "Case")


MOTIONLOGS = ["chain_rectace_rotace_rot", "concurrent_rect_rot", "rotsinu_radians_chain", "rect_wait_rectace_prev",
              "chain_cir_cirace_cirsinu", "rectfile_then_rotfile", "nested_parent_rect_child_rot"]


@pytest.mark.parametrize("name", MOTIONLOGS)
def test_motion_laws_other_movements_match_reference_log(name):
    """Accelerated movements, chains (next), a delayed start, a wait inside a chain, two concurrent movements of one object (translation
    composed with rotation), radians, an event with a finish time, circular movements, file-driven movements (positions / angles
    interpolated from data files): tests/golden/motionlog_*.npz hold what JDsMotion computed at each of
    40 steps of a reference run (generator: tests/golden/make_motionlogs.py)."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "motionlog_" + name + ".npz"))
    tdt = np.asarray(z["t_dt"]).reshape(-1, 2)
    nobj = int(z["nobj"])
    val = np.asarray(z["val"]).reshape(tdt.shape[0], nobj, 23)
    with tempfile.TemporaryDirectory() as tmp:
        open(os.path.join(tmp, "Case.xml"), "wb").write(bytes(z["xml"]))
        for k in z.files:
            if k.startswith("file_"):                      # data files of the file-driven movements, next to the case
                open(os.path.join(tmp, k[5:]), "wb").write(bytes(z[k]))
        motion = dsph.CaseMotion(tmp, "Case")
    assert motion.nobj == nobj == 1
    types = set()
    worst = 0.0
    for s in range(tdt.shape[0]):
        objs, active = motion.process(float(tdt[s, 0]), float(tdt[s, 1]))
        d = objs[0]; ref = val[s, 0]
        assert d.type == int(ref[0]), (name, s, d.type, ref[0], float(tdt[s, 0]))
        types.add(d.type)
        if d.type == 1:
            got = np.array(list(d.linmov) + list(d.linvel)); want = ref[1:7]
        elif d.type == 2:
            got = np.array(list(d.matmov)); want = ref[7:23]
        else:
            continue
        err = float(np.abs(got - want).max() / max(1e-300, np.abs(want).max()))
        worst = max(worst, err)
        assert err < 1e-11, (name, s, got, want)
    print(name, "types seen:", sorted(types), "worst relative difference: %.2e" % worst)
    motion.close()
