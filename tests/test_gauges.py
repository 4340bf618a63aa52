"""In-solver gauges pinned to the REFERENCE: tests/golden/*_gauges.npz hold the results of the reference's own
JGaugeVelocity / JGaugeSwl / JGaugeMaxZ / JGaugeForce (CPU, S/JDsGaugeItem.cpp:325-367,575-600,736-776,986-1041) dumped by
oracle/ref_harness.cpp at the initial state and at selected steps of a dam break whose Case.xml carries a <special><gauges> block."""
import os

import numpy as np
import pytest

import parity
from parity import Golden
import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import capi

GAUGE_CASES = ["db_sym_wend_art_ddt2_gauges", "db_ver_wend_art_ddt2_gauges"]


def _write_case(g, tmp_path):
    """Case.xml (with the gauges block, as the reference read it) + Case.bi4 of the golden."""
    z = g.z
    cd = dsph.CaseData(g.cfg, z["case/idp"], z["case/pos"], z["case/velrhop"], int(z["case/npb"]), "Case",
                       list(z["case/caseposmin"]), list(z["case/caseposmax"]))
    dsph.save_case(cd, str(tmp_path), "Case")
    open(os.path.join(str(tmp_path), "Case.xml"), "wb").write(z["case/xml"].tobytes())
    return cd


def test_gauge_definitions_are_read_from_case_xml(tmp_path):
    g = Golden(GAUGE_CASES[0])
    cd = _write_case(g, tmp_path)
    gs = dsph.case_gauges(str(tmp_path), "Case")
    assert [x.name.decode() for x in gs] == ["Vcol", "Vtoe", "Vdry", "Hcol", "Front", "Zcol", "Zdry", "Ftank"]
    assert [x.type for x in gs] == [1, 1, 1, 2, 2, 3, 3, 4]
    assert all(x.computedt == 0 and x.output == 1 and x.computeend == 10 for x in gs)
    assert list(gs[0].point0) == [0.2, 0.33, 0.1] and list(gs[4].point2) == [1.55, 0.33, 0.03]
    dp, h, mf = cd.cfg.dp, float(cd.cfg.kernelh), float(cd.cfg.massfluid)
    assert abs(gs[3].pointdp - 0.5 * dp) < 1e-7 * dp and abs(gs[3].masslimit - 0.5 * mf) <= 1e-6 * mf and abs(gs[4].masslimit - 0.5 * mf) <= 1e-6 * mf
    assert abs(gs[5].distlimit - 3 * dp) <= 1e-6 * dp and abs(gs[6].distlimit - 2 * h) <= 1e-6 * h and gs[5].height == pytest.approx(0.4)
    assert (gs[7].mkbound, gs[7].idbegin, gs[7].count, gs[7].code) == (0, 0, cd.npb, capi.CODE_TYPE_FIXED)


def _check(res, ref, cfg, what):
    """Tolerances: 1e-5 of the natural scale of each quantity (north_star's single-step tolerance): velocities against
    sqrt(g*H0), positions against the column height H0, the force against the hydrostatic force scale rho0*g*H0*A (A = 0.4 x 0.67 m)."""
    H0 = 0.3
    vs, ps, fs = (9.81 * H0) ** 0.5, H0, float(cfg.rhopzero) * 9.81 * H0 * 0.4 * 0.67
    rep = {}
    for i, (r, t) in enumerate(zip(res, ref["type"])):
        v = ref["values"][i]
        assert r.computed == 1 and abs(r.timestep - float(ref["time"][i])) <= 1e-6 * max(float(ref["time"][i]), 1e-3), (what, i, r.timestep, ref["time"][i])   # simulation time after up to 60 chained steps
        got = np.array(list(r.value), np.float64)
        if t == 0:
            err = np.abs(got - v[3:6]).max() / vs
        elif t == 1:
            err = np.abs(got - v[6:9]).max() / ps
        elif t == 2:
            err = abs(got[2] - v[5]) / ps
        else:
            err = np.abs(got - v[0:3]).max() / fs
        rep[i] = float(err)
        assert err <= 1e-5, (what, "gauge", i, "type", int(t), got, v, err)
    return rep


def _run_gauges(lib, name, tmp_path):
    g = Golden(name)
    _write_case(g, tmp_path)
    gauges = dsph.case_gauges(str(tmp_path), "Case")
    ctx = parity.make_ctx(lib, g.cfg, 20000)
    z = g.z
    ctx.upload(z["case/idp"], z["case/pos"], z["case/velrhop"], int(z["case/npb"]), z["case/code"])
    ctx.celldiv(True)
    ctx.gauges_config(gauges)
    # initial state: RunGaugeSystem(TimeStep=0) (S/JSphCpuSingle.cpp:1070)
    if g.has(0, "ginit") or ("s0000_ginit/count" in z.files):
        print(name, "init", _check(ctx.gauges_run(0.0), g.stage(0, "ginit"), g.cfg, "init"))
    ctx.gauges_set_time(0.0)
    for step in range(g.nsteps):
        ctx.step()                       # runs the gauges after the update, before the closing divide (S/JSphCpuSingle.cpp:1094-1098)
        if step in g.steps:
            print(name, step, _check(ctx.gauges_results(), g.stage(step, "gauges"), g.cfg, "step %d" % step))
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", GAUGE_CASES)
def test_gpu_gauges_match_reference(gpu_lib, name, tmp_path):
    _run_gauges(gpu_lib, name, tmp_path)


@pytest.mark.parametrize("name", GAUGE_CASES)
def test_oracle_gauges_match_reference(oracle_lib, name, tmp_path):
    """The CPU restatement of the four gauges (oracle/sph_oracle.cpp: Oracle::GaugesRun) against the same reference results, no GPU."""
    _run_gauges(oracle_lib, name, tmp_path)


@pytest.mark.gpu
def test_gpu_driver_writes_gauge_csv_files(gpu_lib, tmp_path):
    """The stand-alone driver reads <special><gauges> and writes Gauges<Type>_<name>.csv with the reference's columns
    (S/JDsGaugeItem.cpp:276-298 ...); the rows of the steps the golden holds equal the reference's results to the 6 digits of %g."""
    g = Golden(GAUGE_CASES[0])
    _write_case(g, tmp_path)
    out = tmp_path / "out"
    dsph.run_case(str(tmp_path), "Case", str(out), nsteps=g.nsteps, savepart=False)
    names = sorted(f for f in os.listdir(str(out)) if f.startswith("Gauges"))
    assert names == ["GaugesForce_Ftank.csv", "GaugesMaxZ_Zcol.csv", "GaugesMaxZ_Zdry.csv", "GaugesSWL_Front.csv", "GaugesSWL_Hcol.csv",
                     "GaugesVel_Vcol.csv", "GaugesVel_Vdry.csv", "GaugesVel_Vtoe.csv"], names
    head = open(str(out / "GaugesSWL_Front.csv")).readline().strip()
    assert head == "time [s];swlx [m];swly [m];swlz [m];pos0x [m];pos0y [m];pos0z [m];pos2x [m];pos2y [m];pos2z [m]"
    rows = np.genfromtxt(str(out / "GaugesSWL_Front.csv"), delimiter=";", skip_header=1)
    assert rows.shape == (g.nsteps + 1, 10)                         # initial state + every step (computedt = outputdt = 0)
    for step in g.steps:
        ref = g.stage(step, "gauges")
        assert abs(rows[step + 1, 0] - float(ref["time"][4])) <= 2e-6 * float(ref["time"][4])
        assert abs(rows[step + 1, 1] - ref["values"][4][6]) <= 3e-6
    f = np.genfromtxt(str(out / "GaugesForce_Ftank.csv"), delimiter=";", skip_header=1)
    ref = g.stage(g.steps[-1], "gauges")["values"][7]
    assert np.abs(f[g.steps[-1] + 1, 2:5] - ref[0:3]).max() <= 2e-5 * np.abs(ref[0:3]).max()
