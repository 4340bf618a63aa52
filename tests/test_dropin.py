"""The compiled drop-in: oracle/_ref/DualSPHysics5.2_b200dropin is the REFERENCE's own GPU program (main.cpp, JSph,
JSphGpu, JSphGpuSingle ... compiled from the unmodified sources by `make -C oracle gpu52_b200`) with eight member functions rebound to
libdsph_b200.so by integration/JSphGpuSingle_b200.cpp.  On C1 (BASELINE.json configs[0], ~170 k particles) it must write the PART files
the stock reference writes: same particles, same output instants, state within the tolerances of the 1-s validation."""
import os
import re
import subprocess

import numpy as np
import pytest

import dualsphysics_multilayer_b200 as dsph

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "oracle", "_ref")
DROPIN = os.path.join(REF, "DualSPHysics5.2_b200dropin")
STOCKGPU = os.path.join(REF, "DualSPHysics5.2_linux64")
STOCKCPU = os.path.join(REF, "DualSPHysics5.2CPU_linux64")
MKCASE = os.path.join(ROOT, "dualsphysics_multilayer_b200", "bin", "mkcase")


def test_dropin_binary_is_the_reference_program_bound_to_the_library():
    """CPU check: the binary exists where the reference was available, links libdsph_b200.so, and its eight rebound member functions
    are the binding's (strong) definitions."""
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/DualSPHysics5.2_b200dropin not built here (make -C oracle gpu52_b200 needs /root/reference)")
    ldd = subprocess.run(["ldd", DROPIN], stdout=subprocess.PIPE, text=True).stdout
    assert "libdsph_b200.so" in ldd and "not found" not in ldd.split("libdsph_b200.so")[1].split("\n")[0], ldd
    nm = subprocess.run(["nm", "-C", DROPIN], stdout=subprocess.PIPE, text=True).stdout
    for sym in ("JSphGpuSingle::RunCellDivide(bool)", "JSphGpuSingle::Interaction_Forces(TpInterStep)", "JSphGpuSingle::ComputeStep_Sym()",
                "JSphGpuSingle::ComputeStep_Ver()", "JSphGpuSingle::RunGaugeSystem(double, bool)",
                "JSphGpu::ParticlesDataDown(unsigned int, unsigned int, bool, bool)", "JCellDivGpu::GetDomainLimits(bool, unsigned int) const",
                "JSphGpu::RunMotion(double)"):
        assert re.search(r" T " + re.escape(sym) + r"$", nm, flags=re.M), sym
    assert re.search(r" U dsph_forces_run", nm) and re.search(r" U dsph_celldiv_run", nm) and re.search(r" U dsph_motion_run", nm)
    assert " T JSphGpuSingle::Run(" in nm and " T JSph::SaveData(" in nm          # the reference's own driver and output code


def _run(binary, tmp, out, extra):
    r = subprocess.run([binary, "Case", out] + extra, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "Finished execution (code=0)" in r.stdout, r.stdout[-3000:]
    return r.stdout


def _state(dirname, part):
    cd, t, step = dsph.load_part(os.path.join(dirname, "Part_%04d.bi4" % part))
    o = np.argsort(cd.idp)
    return cd.idp[o], cd.pos[o], cd.velrhop[o], t, step


@pytest.mark.gpu
def test_gpu_dropin_writes_the_reference_part_files_on_c1(gpu_lib, tmp_path):
    if not (os.path.exists(DROPIN) and os.path.exists(STOCKCPU)):
        pytest.skip("drop-in or reference CPU binary not built")
    tmp = str(tmp_path)
    subprocess.check_call([MKCASE, "dambreak", "170000", tmp, "Case"], stdout=subprocess.DEVNULL)
    opts = ["-svres:1", "-tmax:0.02", "-tout:0.01"]
    txt = _run(DROPIN, tmp, "out_dropin", ["-gpu"] + opts)
    assert "libdsph_b200" in txt
    _run(STOCKCPU, tmp, "out_cpu", ["-cpu", "-ompthreads:%d" % min(os.cpu_count() or 1, 32)] + opts)
    have_gpu_ref = os.path.exists(STOCKGPU)
    if have_gpu_ref:
        _run(STOCKGPU, tmp, "out_gpu", ["-gpu"] + opts)
    rep = {}
    for part in (0, 1, 2):
        ia, pa, va, ta, sa = _state(os.path.join(tmp, "out_dropin", "data") if os.path.isdir(os.path.join(tmp, "out_dropin", "data")) else os.path.join(tmp, "out_dropin"), part)
        ib, pb, vb, tb, sb = _state(os.path.join(tmp, "out_cpu", "data") if os.path.isdir(os.path.join(tmp, "out_cpu", "data")) else os.path.join(tmp, "out_cpu"), part)
        assert np.array_equal(ia, ib), "same particles in the PART files"
        assert sa == sb and abs(ta - tb) <= 1e-7 * max(tb, 1e-3), (part, sa, sb, ta, tb)
        dpos = float(np.abs(pa - pb).max())
        dvel = float(np.linalg.norm((va - vb).astype(np.float64)) / np.linalg.norm(vb.astype(np.float64)))
        rep[part] = dict(step=sa, time=ta, dpos=dpos, dvelrhop=dvel)
        if have_gpu_ref:
            ic, pc, vc, tc, sc = _state(os.path.join(tmp, "out_gpu", "data") if os.path.isdir(os.path.join(tmp, "out_gpu", "data")) else os.path.join(tmp, "out_gpu"), part)
            rep[part]["stock_gpu_dpos"] = float(np.abs(pc - pb).max()) if np.array_equal(ic, ib) else -1.0
        # single-precision trajectories of ~110 steps each: far inside the 1 % of the 1-s validation (dp = 8.5 mm)
        assert dpos <= 1e-6 and dvel <= 1e-5, rep
    print("drop-in vs reference CPU solver on C1:", rep)
    runout = open(os.path.join(tmp, "out_dropin", "Run.out")).read()
    assert "Particles of simulation (initial): " in runout and "NL-SortData" in runout


@pytest.mark.gpu
def test_gpu_dropin_runs_a_moving_boundary_case_like_the_reference(gpu_lib, tmp_path):
    """Moving boundaries through the drop-in: the reference's own JDsMotion evaluates <motion>, JSphGpu::RunMotion (rebound) hands the
    StMotionData to dsph_motion_run.  An oscillating piston next to the water column, 60 k particles, against the reference CPU solver."""
    if not (os.path.exists(DROPIN) and os.path.exists(STOCKCPU)):
        pytest.skip("drop-in or reference CPU binary not built")
    tmp = str(tmp_path)
    subprocess.check_call([MKCASE, "movpiston", "60000", tmp, "Case"], stdout=subprocess.DEVNULL)
    opts = ["-svres:1", "-tmax:0.01", "-tout:0.005"]
    _run(DROPIN, tmp, "out_dropin", ["-gpu"] + opts)
    _run(STOCKCPU, tmp, "out_cpu", ["-cpu", "-ompthreads:%d" % min(os.cpu_count() or 1, 32)] + opts)
    cd = dsph.load_case(tmp, "Case")
    nf, nm = int(cd.cfg.casenfixed), int(cd.cfg.casenmoving)
    rep = {}
    for part in (1, 2):
        da = os.path.join(tmp, "out_dropin", "data"); db = os.path.join(tmp, "out_cpu", "data")
        ia, pa, va, ta, sa = _state(da if os.path.isdir(da) else os.path.join(tmp, "out_dropin"), part)
        ib, pb, vb, tb, sb = _state(db if os.path.isdir(db) else os.path.join(tmp, "out_cpu"), part)
        assert np.array_equal(ia, ib) and sa == sb and abs(ta - tb) <= 1e-7 * max(tb, 1e-3), (part, sa, sb, ta, tb)
        plate = float(np.abs(pa[nf:nf + nm] - pb[nf:nf + nm]).max())
        moved = float(np.abs(pa[nf:nf + nm, 0] - cd.pos[np.argsort(cd.idp)][nf:nf + nm, 0]).max())
        dpos = float(np.abs(pa - pb).max())
        dvel = float(np.linalg.norm((va - vb).astype(np.float64)) / np.linalg.norm(vb.astype(np.float64)))
        rep[part] = dict(step=sa, plate_moved=moved, plate_diff=plate, dpos=dpos, dvelrhop=dvel)
        # the plate follows the same law at instants that differ like the variable dt of the two runs does (1e-6 relative)
        assert moved > 1e-4 and plate <= 1e-9 and dpos <= 1e-6 and dvel <= 1e-5, rep
    print("drop-in vs reference CPU solver, moving piston:", rep)


@pytest.mark.gpu
def test_gpu_dropin_runs_an_mdbc_case_like_the_reference(gpu_lib, tmp_path):
    """mDBC through the drop-in: the reference's own LoadBoundNormals / ConfigBoundNormals prepare the vectors particle -> ghost node, the
    binding hands them to dsph_particles_set_boundnormals, and the correction runs inside dsph_forces_run.  60 k particles, Boundary=2 with
    DDT1 (the boundary rule of the density diffusion is on the path), against the reference CPU solver."""
    if not (os.path.exists(DROPIN) and os.path.exists(STOCKCPU)):
        pytest.skip("drop-in or reference CPU binary not built")
    tmp = str(tmp_path)
    subprocess.check_call([MKCASE, "dambreak", "60000", tmp, "Case", "boundary=2", "ddt=1"], stdout=subprocess.DEVNULL)
    opts = ["-svres:1", "-tmax:0.01", "-tout:0.005"]
    txt = _run(DROPIN, tmp, "out_dropin", ["-gpu"] + opts)
    assert 'Boundary="mDBC"' in txt and "libdsph_b200" in txt
    _run(STOCKCPU, tmp, "out_cpu", ["-cpu", "-ompthreads:%d" % min(os.cpu_count() or 1, 32)] + opts)
    rep = {}
    for part in (1, 2):
        da = os.path.join(tmp, "out_dropin", "data"); db = os.path.join(tmp, "out_cpu", "data")
        ia, pa, va, ta, sa = _state(da if os.path.isdir(da) else os.path.join(tmp, "out_dropin"), part)
        ib, pb, vb, tb, sb = _state(db if os.path.isdir(db) else os.path.join(tmp, "out_cpu"), part)
        assert np.array_equal(ia, ib) and sa == sb and abs(ta - tb) <= 1e-7 * max(tb, 1e-3), (part, sa, sb, ta, tb)
        dpos = float(np.abs(pa - pb).max())
        dvel = float(np.linalg.norm((va - vb).astype(np.float64)) / np.linalg.norm(vb.astype(np.float64)))
        rep[part] = dict(step=sa, dpos=dpos, dvelrhop=dvel)
        assert dpos <= 1e-6 and dvel <= 1e-5, rep
    print("drop-in vs reference CPU solver, mDBC:", rep)
