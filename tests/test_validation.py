"""1-s dam-break validation (north_star: gauge pressures and surge-front position over 1 s within 1 %).

Golden series: the reference's stock CPU solver on the 20 k-particle synthetic dam break, measured every 0.05 s
(tests/golden/make_validation.py -> tests/golden/val_dambreak_1s.npz).  What two CPU builds of the same algorithm can agree
on was measured first (reference vs the checker in oracle/, 5 546 steps): front 7e-4, column height 2e-4, bed gauges under
the column G1 1.2e-3 and G2 8.8e-3; the gauge in the thin downstream tongue (G3) and the end-wall impact gauge (G4) are
chaotic at the particle level (the two CPU builds differ by 24 Pa and 360 Pa there), so they are checked against the
hydrostatic pressure scale rho0*g*H0 instead of their own (near-zero / spiky) value."""
import os
import sys

import numpy as np
import pytest

import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200.capi import SphContext

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import measure  # noqa: E402
from make_validation import gauges_for, series_from_parts  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "val_dambreak_1s.npz")
COL = dict(time=0, step=1, np=2, front=3, height=4, g1=5, g2=6, g3=7, g4=8)


def _case(tmp_path, z):
    cd = dsph.make_case("dambreak", int(z["np_target"]))
    dsph.save_case(cd, str(tmp_path), "Case")
    cd = dsph.load_case(str(tmp_path), "Case")
    assert bytes(cd.cfg) == z["config"].tobytes(), "the generated case must be the one the golden series was made with"
    return cd


def _compare(ours, ref, upto):
    """Relative differences after interpolating the reference series onto our output times."""
    ours = ours[(ours[:, 0] > 0) & (ours[:, 0] <= upto + 1e-3)]
    t = ours[:, 0]
    out = {}
    for k in ("front", "height", "g1", "g2", "g3", "g4"):
        r = np.interp(t, ref[:, 0], ref[:, COL[k]])
        out[k] = (ours[:, COL[k]], r)
    return t, out


def test_oracle_tracks_reference_for_a_tenth_of_a_second(oracle_lib, tmp_path):
    """CPU-only: the checker, stepped through the C ABI, against the golden series (532 steps)."""
    z = np.load(GOLD)
    cd = _case(tmp_path, z)
    ctx = SphContext(oracle_lib, cd.cfg, 0)
    ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, cd.codes())
    ctx.celldiv(True)
    t, nstep, nextt, rows = 0.0, 0, 0.05, []
    while t < 0.1:
        dt, _ = ctx.step(); t += dt; nstep += 1
        if t >= nextt:
            st = ctx.download()
            m = measure.measure_state(st["pos"], st["velrhop"][:, 3], st["idp"], cd.cfg, gauges_for(cd))
            rows.append((t, nstep, ctx.np_, m["front"], m["height"], *m["press"]))
            nextt += 0.05
    ctx.close()
    tt, cmp = _compare(np.array(rows), z["series"], 0.1)
    assert len(tt) == 2
    for k in ("front", "height", "g1", "g2"):
        a, r = cmp[k]
        assert np.all(np.abs(a - r) <= 1e-3 * np.abs(r)), (k, a, r)


@pytest.mark.gpu
def test_gpu_dambreak_1s_front_and_gauges_within_1_percent(gpu_lib, tmp_path):
    """The stand-alone driver (dsph_run_case: Case.xml+bi4 in, Part_NNNN.bi4 out) on the GPU for 1 s of physical time."""
    z = np.load(GOLD)
    cd = _case(tmp_path, z)
    out = tmp_path / "out"
    np.savetxt(str(tmp_path / "Case_Points.txt"), np.array(gauges_for(cd)), header="probe points (x y z)")   # driver: device-side probes
    dsph.run_case(str(tmp_path), "Case", str(out), device=0, nsteps=0, savepart=True)
    ours = series_from_parts(str(out), cd)
    # the driver's own probe output (dsph_measure_points kernel) must agree with the numpy measurement of its PART files
    csv = np.genfromtxt(str(out / "MeasurePoints.csv"), delimiter=";", skip_header=1)
    pcsv = csv[:, 6].reshape(-1, 4)
    assert pcsv.shape[0] == len(ours)
    scale0 = float(cd.cfg.rhopzero) * 9.81 * 0.3
    assert np.max(np.abs(pcsv - ours[:, 5:9])) <= 2e-4 * scale0, np.max(np.abs(pcsv - ours[:, 5:9]))
    ref = z["series"]
    assert len(ours) == len(ref) == 21, (len(ours), len(ref))
    assert np.all(ours[:, COL["np"]] == ref[:, COL["np"]]), "no particle may be lost that the reference keeps"
    t, cmp = _compare(ours, ref, 1.0)
    rep = {}
    for k in ("front", "height", "g1", "g2"):
        a, r = cmp[k]
        rep[k] = float(np.max(np.abs(a - r) / np.abs(r)))
    scale = float(cd.cfg.rhopzero) * 9.81 * 0.3          # hydrostatic pressure of the initial column
    for k in ("g3", "g4"):
        a, r = cmp[k]
        rep[k + "_abs_over_rho_g_H0"] = float(np.max(np.abs(a - r)) / scale)
    print("1-s validation, max relative difference to the reference CPU solver:", rep)
    assert rep["front"] <= 0.01 and rep["height"] <= 0.01, rep
    assert rep["g1"] <= 0.01 and rep["g2"] <= 0.01, rep
    assert rep["g3_abs_over_rho_g_H0"] <= 0.05, rep       # thin tongue: a handful of particles under the kernel


@pytest.mark.gpu
def test_gpu_restart_from_part_continues_the_run(gpu_lib, tmp_path):
    """-partbegin: a run restarted from Part_0002 of a previous run ends where the uninterrupted run ends (the restart rebuilds
    the cell order from the id order of the file, so the force sums are reordered: single-step tolerance, not bit equality)."""
    cd = dsph.make_case("dambreak", 30000)
    dsph.save_case(cd, str(tmp_path), "Case")
    full = tmp_path / "full"; rest = tmp_path / "rest"
    dsph.run_case(str(tmp_path), "Case", str(full), timemax=0.02, timeout=0.005)
    dsph.run_case(str(tmp_path), "Case", str(rest), timemax=0.02, timeout=0.005, partbegin=2, partbegindir=str(full))
    a, ta, sa = dsph.load_part(str(full / "Part_0004.bi4"))
    b, tb, sb = dsph.load_part(str(rest / "Part_0004.bi4"))
    assert not (rest / "Part_0002.bi4").exists() and (rest / "Part_0003.bi4").exists()
    assert a.np_ == b.np_ and abs(ta - tb) <= 1e-9 * ta
    oa = np.argsort(a.idp); ob = np.argsort(b.idp)
    assert np.array_equal(a.idp[oa], b.idp[ob])
    assert np.abs(a.pos[oa] - b.pos[ob]).max() < 1e-7
    den = np.linalg.norm(a.velrhop[oa].astype(np.float64))
    assert np.linalg.norm((a.velrhop[oa] - b.velrhop[ob]).astype(np.float64)) / den < 1e-5


def _sorted_state(cd):
    o = np.argsort(cd.idp)
    return cd.idp[o], cd.pos[o], cd.velrhop[o]


@pytest.mark.gpu
def test_gpu_driver_output_set_pieces_and_excluded_particles(gpu_lib, tmp_path):
    """the driver's asynchronous PART pipeline writes Part_Head.ibi4, PartInfo.ibi4, PartOut_000.obi4 and, with a piece
    limit below the particle count, Part_pXX_NNNN.bi4 pieces; the pieces hold the state of the single-file run bit for bit, and the
    particles the divide excluded (a RhopOut window the flow leaves) are in PartOut with their motive."""
    from test_partset import read_bi4
    cd = dsph.make_case("dambreak", 30000)
    cd.cfg.rhopoutmin, cd.cfg.rhopoutmax = 995.0, float(cd.velrhop[:, 3].max()) + 0.6      # a few particles pass it during the collapse
    dsph.save_case(cd, str(tmp_path), "Case")
    one = tmp_path / "one"; pcs = tmp_path / "pieces"
    dsph.run_case(str(tmp_path), "Case", str(one), timemax=0.03, timeout=0.005)
    dsph.run_case(str(tmp_path), "Case", str(pcs), timemax=0.03, timeout=0.005, piecemax=cd.np_ // 3 + 1)
    names = sorted(os.listdir(str(one)))
    assert {"Part_Head.ibi4", "PartInfo.ibi4", "PartOut_000.obi4", "Part_0000.bi4", "Part_0006.bi4", "Run.out", "Run.csv"} <= set(names), names
    assert {"Part_p00_0006.bi4", "Part_p01_0006.bi4", "Part_p02_0006.bi4", "PartInfo_p02.ibi4"} <= set(os.listdir(str(pcs)))
    a, ta, sa, npa = dsph.load_partset(str(one), 6)
    b, tb, sb, npb_ = dsph.load_partset(str(pcs), 6)
    assert (npa, npb_) == (1, 3) and ta == tb and sa == sb
    for x, y in zip(_sorted_state(a), _sorted_state(b)):
        assert np.array_equal(x, y)
    inf = read_bi4(str(one / "PartInfo.ibi4"), "JPartDataBi4_Info")
    assert [i["name"] for i in inf] == ["JPartDataBi4"] + ["PART_%04d" % k for k in range(7)]
    assert inf[-1]["values"]["Npok"] == a.np_ and inf[-1]["values"]["dtmin"] > 0 and inf[-1]["values"]["dtmax"] >= inf[-1]["values"]["dtmin"]
    po = read_bi4(str(one / "PartOut_000.obi4"), "JPartOutBi4")
    nout = sum(i["values"]["Nout"] for i in po[1:])
    txt = open(str(one / "Run.out")).read()
    print(txt[-900:])
    assert nout > 0 and nout == cd.np_ - a.np_ and ("excluded=%d" % nout) in txt
    mot = np.concatenate([i["arrays"]["Motive"] for i in po[1:]])
    rh = np.concatenate([i["arrays"]["Rhop"] for i in po[1:]])
    assert np.all(mot == 2) and np.all((rh < 995.0) | (rh > cd.cfg.rhopoutmax))
    assert "PART pipeline" in txt


@pytest.mark.gpu
def test_gpu_driver_slab_run_matches_single_gpu_run(gpu_lib, tmp_path):
    """bin/dsph_b200 -mgpu:2 (one host thread, one NCCL rank and one PART piece per GPU) against the single-GPU run of
    the same case: same particle set in the pieces, same time, state within the single-step tolerance."""
    if dsph.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cd = dsph.make_case("dambreak", 200000)
    dsph.save_case(cd, str(tmp_path), "Case")
    one = tmp_path / "one"; two = tmp_path / "two"
    dsph.run_case(str(tmp_path), "Case", str(one), nsteps=30, timeout=2e-4)
    dsph.run_case(str(tmp_path), "Case", str(two), nsteps=30, timeout=2e-4, mgpu=2, rebalancesteps=10)
    last = max(int(f[5:9]) for f in os.listdir(str(one)) if f.startswith("Part_0") and f.endswith(".bi4"))
    assert last >= 2 and os.path.exists(str(two / ("Part_p01_%04d.bi4" % last))) and os.path.exists(str(two / "Run_p01.out"))
    a, ta, sa, npa = dsph.load_partset(str(one), last)
    b, tb, sb, npb_ = dsph.load_partset(str(two), last)
    assert (npa, npb_) == (1, 2) and sa == sb and abs(ta - tb) <= 1e-9 * ta
    ia, pa, va = _sorted_state(a); ib, pb, vb = _sorted_state(b)
    assert np.array_equal(ia, ib)
    assert np.abs(pa - pb).max() < 1e-7
    assert np.linalg.norm((va - vb).astype(np.float64)) / np.linalg.norm(va.astype(np.float64)) < 1e-5
    print(open(str(two / "Run.out")).read()[-1200:])


def test_mkcase_writes_mdbc_normals_the_reference_accepts(tmp_path):
    """CPU: `mkcase ... boundary=2` writes <case>_Normals.nbi4 (JPartNormalData format); the library reads it back, and the reference CPU
    solver runs the case with Boundary="mDBC" from that file."""
    import subprocess
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    mkcase = os.path.join(root, "dualsphysics_multilayer_b200", "bin", "mkcase")
    subprocess.check_call([mkcase, "dambreak", "3000", str(tmp_path), "Case", "boundary=2", "ddt=1"], stdout=subprocess.DEVNULL)
    assert os.path.exists(str(tmp_path / "Case_Normals.nbi4"))
    cd = dsph.load_case(str(tmp_path), "Case")
    assert cd.cfg.tboundary == 2
    n = dsph.case_normals(str(tmp_path), "Case")
    assert n.shape == (cd.npb, 3)
    hd = 0.5 * cd.cfg.dp
    assert np.all(np.isin(np.round(n / hd).astype(int), (-1, 0, 1))) and np.all(np.abs(n).sum(axis=1) > 0)
    bed = np.isclose(cd.pos[:cd.npb, 2], 0.0)
    assert np.allclose(n[bed, 2], hd) and np.all(n[~bed, 2] == 0)
    ref = os.path.join(root, "oracle", "_ref", "DualSPHysics5.2CPU_linux64")
    if os.path.exists(ref):
        r = subprocess.run([ref, "Case", "out", "-cpu", "-ompthreads:2", "-sv:none", "-nsteps:3"], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert 'Boundary="mDBC"' in r.stdout and 'NormalDataFile="Case_Normals.nbi4"' in r.stdout and "Finished execution (code=0)" in r.stdout, r.stdout[-2000:]


@pytest.mark.gpu
def test_gpu_driver_runs_mdbc_case_like_the_reference(gpu_lib, tmp_path):
    """bin/dsph_b200 on an mDBC case (Boundary=2 in Case.xml + Case_Normals.nbi4): its last PART equals the final state the reference CPU
    solver reached after the same number of steps (golden dump db_sym_wend_art_ddt1_mdbc)."""
    import subprocess
    from parity import Golden
    g = Golden("db_sym_wend_art_ddt1_mdbc")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    mkcase = os.path.join(root, "dualsphysics_multilayer_b200", "bin", "mkcase")
    subprocess.check_call([mkcase, "dambreak", "3000", str(tmp_path), "Case", "tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "boundary=2"], stdout=subprocess.DEVNULL)
    cd = dsph.load_case(str(tmp_path), "Case")
    assert np.array_equal(cd.pos, g.z["case/pos"]) and cd.cfg.tboundary == 2
    out = tmp_path / "out"
    dsph.run_case(str(tmp_path), "Case", str(out), nsteps=g.nsteps, timeout=1e-5)          # a PART at every step: the last one is the state after nsteps
    assert "mDBC" in open(str(out / "Run.out")).read()
    a, ta, sa, _ = dsph.load_partset(str(out), g.nsteps)
    fin = g.stage(g.nsteps, "final")
    assert sa == g.nsteps and abs(ta - float(fin["timestep"])) <= 1e-6 * ta
    oa = np.argsort(a.idp); ob = np.argsort(fin["idp"])
    assert np.array_equal(a.idp[oa], fin["idp"][ob])
    assert np.abs(a.pos[oa] - fin["pos"][ob]).max() < 1e-6
    assert np.linalg.norm((a.velrhop[oa] - fin["velrhop"][ob]).astype(np.float64)) / np.linalg.norm(fin["velrhop"][ob].astype(np.float64)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind,opts", [("mv_sym_gate_rect_wait", "movgate", ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2"]),
                                            ("mv_ver_flap_rotsinu", "movflap", ["tstep=1", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "verletsteps=10"])])
def test_gpu_driver_runs_moving_boundary_case_like_the_reference(gpu_lib, tmp_path, name, kind, opts):
    """bin/dsph_b200 on a case with a moving object (<moving> block + <motion> laws in Case.xml): its last PART equals the final state the
    reference CPU solver reached after the same number of steps (golden dumps mv_*), moved plate included."""
    import subprocess
    from parity import Golden
    g = Golden(name)
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    mkcase = os.path.join(root, "dualsphysics_multilayer_b200", "bin", "mkcase")
    subprocess.check_call([mkcase, kind, "3000", str(tmp_path), "Case"] + opts, stdout=subprocess.DEVNULL)
    cd = dsph.load_case(str(tmp_path), "Case")
    assert np.array_equal(cd.pos, g.z["case/pos"]) and int(cd.cfg.casenmoving) > 0
    out = tmp_path / "out"
    dsph.run_case(str(tmp_path), "Case", str(out), nsteps=g.nsteps, timeout=1e-5)
    assert "Motion: 1 moving object" in open(str(out / "Run.out")).read()
    a, ta, sa, _ = dsph.load_partset(str(out), g.nsteps)
    fin = g.stage(g.nsteps, "final")
    assert sa == g.nsteps and abs(ta - float(fin["timestep"])) <= 1e-6 * ta
    oa = np.argsort(a.idp); ob = np.argsort(fin["idp"])
    assert np.array_equal(a.idp[oa], fin["idp"][ob])
    nf, nm = int(cd.cfg.casenfixed), int(cd.cfg.casenmoving)
    moved = np.abs(a.pos[oa][nf:nf + nm] - cd.pos[np.argsort(cd.idp)][nf:nf + nm]).max()
    assert moved > 1e-4, moved                                        # the plate did move
    # ... as in the reference run: the library's own evaluation of the motion laws agrees with JDsMotion to ~1e-13 per step (last bits of
    # sin / cos and of the matrix products), accumulated over the steps
    assert np.abs(a.pos[oa][nf:nf + nm] - fin["pos"][ob][nf:nf + nm]).max() < 1e-10
    assert np.abs(a.pos[oa] - fin["pos"][ob]).max() < 1e-6
    assert np.linalg.norm((a.velrhop[oa] - fin["velrhop"][ob]).astype(np.float64)) / np.linalg.norm(fin["velrhop"][ob].astype(np.float64)) < 1e-5


@pytest.mark.gpu
def test_gpu_driver_slab_run_with_moving_piston_matches_single_gpu_run(gpu_lib, tmp_path):
    """A moving object in an x-slab run: every rank evaluates the motion laws and moves the plate particles it owns; halo copies of the
    plate are refreshed by the exchange.  -mgpu:2 against the single-GPU run of the same case."""
    import subprocess
    if dsph.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    mkcase = os.path.join(root, "dualsphysics_multilayer_b200", "bin", "mkcase")
    subprocess.check_call([mkcase, "movpiston", "200000", str(tmp_path), "Case"], stdout=subprocess.DEVNULL)
    one = tmp_path / "one"; two = tmp_path / "two"
    dsph.run_case(str(tmp_path), "Case", str(one), nsteps=30, timeout=2e-4)
    dsph.run_case(str(tmp_path), "Case", str(two), nsteps=30, timeout=2e-4, mgpu=2)
    last = max(int(f[5:9]) for f in os.listdir(str(one)) if f.startswith("Part_0") and f.endswith(".bi4"))
    a, ta, sa, npa = dsph.load_partset(str(one), last)
    b, tb, sb, npb_ = dsph.load_partset(str(two), last)
    assert (npa, npb_) == (1, 2) and sa == sb and abs(ta - tb) <= 1e-9 * ta
    ia, pa, va = _sorted_state(a); ib, pb, vb = _sorted_state(b)
    assert np.array_equal(ia, ib)
    assert np.abs(pa - pb).max() < 1e-7
    assert np.linalg.norm((va - vb).astype(np.float64)) / np.linalg.norm(va.astype(np.float64)) < 1e-5


@pytest.mark.gpu
def test_gpu_driver_slab_run_of_the_periodic_box_matches_single_gpu_run(gpu_lib, tmp_path):
    """bin/dsph_b200 -mgpu:2 on the fully periodic box: the two ranks form a ring through the periodic face along x and copy along y and z
    themselves; against the single-GPU run of the same case."""
    if dsph.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cd = dsph.make_case("periodicbox", 200000)
    dsph.save_case(cd, str(tmp_path), "Case")
    one = tmp_path / "one"; two = tmp_path / "two"
    dsph.run_case(str(tmp_path), "Case", str(one), nsteps=20, timeout=1e-3)
    dsph.run_case(str(tmp_path), "Case", str(two), nsteps=20, timeout=1e-3, mgpu=2)
    last = max(int(f[5:9]) for f in os.listdir(str(one)) if f.startswith("Part_0") and f.endswith(".bi4"))
    a, ta, sa, npa = dsph.load_partset(str(one), last)
    b, tb, sb, npb_ = dsph.load_partset(str(two), last)
    assert last >= 1 and (npa, npb_) == (1, 2) and sa == sb and abs(ta - tb) <= 1e-9 * ta
    ia, pa, va = _sorted_state(a); ib, pb, vb = _sorted_state(b)
    assert np.array_equal(ia, ib)
    d = np.abs(pa - pb); d = np.minimum(d, np.abs(d - 1.0))          # a particle may sit on either side of a periodic face (period 1)
    assert d.max() < 1e-7
    assert np.linalg.norm((va - vb).astype(np.float64)) / np.linalg.norm(va.astype(np.float64)) < 1e-5


@pytest.mark.gpu
def test_gpu_driver_slab_run_writes_the_gauges_of_the_single_gpu_run(gpu_lib, tmp_path):
    """In-solver gauges in an x-slab run: every rank sums over the particles it owns, the partial results are all-reduced (sum for velocity,
    line masses and the force, max for maxz) and rank 0 writes the Gauges*.csv files.  -mgpu:2 against the single-GPU run of the same case:
    the same rows at the same instants, values equal up to the order of the float sums."""
    import subprocess, sys
    if dsph.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    sys.path.insert(0, os.path.join(root, "tests", "golden"))
    from make_golden import inject_gauges
    mkcase = os.path.join(root, "dualsphysics_multilayer_b200", "bin", "mkcase")
    subprocess.check_call([mkcase, "dambreak", "200000", str(tmp_path), "Case"], stdout=subprocess.DEVNULL)
    inject_gauges(str(tmp_path / "Case.xml"))
    one = tmp_path / "one"; two = tmp_path / "two"
    dsph.run_case(str(tmp_path), "Case", str(one), nsteps=25, timeout=1e-3)
    dsph.run_case(str(tmp_path), "Case", str(two), nsteps=25, timeout=1e-3, mgpu=2)
    files = sorted(f for f in os.listdir(str(one)) if f.startswith("Gauges") and f.endswith(".csv"))
    assert len(files) == 8, files
    worst = {}
    for f in files:
        a = np.genfromtxt(str(one / f), delimiter=";", skip_header=1); b = np.genfromtxt(str(two / f), delimiter=";", skip_header=1)
        assert a.shape == b.shape and a.shape[0] >= 20, (f, a.shape, b.shape)
        scale = max(1e-3, float(np.abs(a[:, 1:]).max()))
        worst[f] = float(np.abs(a - b).max() / scale)
        assert worst[f] < 2e-5, (f, worst[f])
    print("gauges of the slab run vs single GPU (max difference / largest value):", worst)
