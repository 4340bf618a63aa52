"""CUDA path (libdsph_b200.so through the C ABI) against (a) the golden dumps of the reference solver and (b) the CPU
checker in oracle/ on larger seeded cases.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

import parity
from parity import Golden, rl2
import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import capi
from dualsphysics_multilayer_b200.capi import SphContext

pytestmark = pytest.mark.gpu

# north_star: cell assignment and sorted order bit-exact; single-step Ace/Ar/dt within 1e-5 relative L2 (fp32)
TOL = dict(default=2e-5, ace=1e-5, ar=1e-5, dt=1e-5, viscdt=1e-5, acemax=1e-5, velmax=1e-6, pos=1e-8, state=1e-6)
# ONE tolerance for every case, stage and step.  Round 1 loosened Ace to 1e-4 for the periodic box and did not check Ar at step 0;
# with the pressure of the reorder rounded like the reference's EOS both are inside 1e-5 (measured, tools/parity_floors.py on the 17
# dumps: Ace <= 2.7e-6, Ar <= 6.7e-6, the periodic-box corrector of step 0 at 4.6e-7).


def tol_for(name):
    return TOL


# mDBC: the corrected boundary density is compared through rho-rho0 (the part that carries the pressure); one ulp of a float near 1000 is
# 6e-5 of that deviation, so 1e-5 of its norm means "bit-identical except for a handful of last-bit differences" (measured: <= 7 of 692
# wall particles differ, by 1 ulp: 2.4e-6).
TOL["mdbc_rhop"] = 1e-5


@pytest.mark.parametrize("name", parity.GOLDEN_CASES + parity.MDBC_CASES + parity.MOVING_CASES + parity.MOVING_MDBC_CASES)
def test_gpu_stages_vs_reference(gpu_lib, name):
    g = Golden(name)
    for step in g.steps:
        res = parity.run_step_stages(gpu_lib, g, step, tol_for(name), capacity=20000)
        print(name, step, res)


@pytest.mark.parametrize("name", parity.GOLDEN_CASES + parity.MDBC_CASES + parity.MOVING_CASES + parity.MOVING_MDBC_CASES)
def test_gpu_trajectory_vs_reference(gpu_lib, name):
    g = Golden(name)
    out = parity.run_trajectory(gpu_lib, g, capacity=20000)
    print(name, out)
    assert out["np"][0] == out["np"][1]
    assert out["time"] < 1e-5
    assert out["pos"] < 1e-6 and out["velrhop"] < 1e-4


def _both(gpu_lib, oracle_lib, cd, capacity=0):
    code = cd.codes()
    a = SphContext(gpu_lib, cd.cfg, capacity)
    b = SphContext(oracle_lib, cd.cfg)
    for c in (a, b):
        c.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, code)
    return a, b


@pytest.mark.parametrize("kind,n,nsteps", [("dambreak", 60000, 12), ("periodicbox", 27000, 6), ("nn2layer", 40000, 8),
                                           # tiny inputs: a single partly filled CTA (30 fluid particles), a few CTAs, a fluid-only box of 8^3
                                           ("dambreak", 300, 6), ("dambreak", 2500, 6), ("periodicbox", 343, 4), ("nn2layer", 1500, 4)])
def test_gpu_vs_oracle_steps(gpu_lib, oracle_lib, kind, n, nsteps):
    """Same seeded case, whole steps on both sides: order bit-exact after every divide, forces within tolerance."""
    cd = dsph.make_case(kind, n)
    a, b = _both(gpu_lib, oracle_lib, cd)
    ia, ib = a.celldiv(True), b.celldiv(True)
    for s in range(nsteps):
        sa, sb = a.download(), b.download()
        assert ia.np == ib.np and ia.npb == ib.npb and ia.npbok == ib.npbok
        assert list(ia.cellmin) == list(ib.cellmin) and list(ia.cellmax) == list(ib.cellmax)
        assert np.array_equal(sa["idp"], sb["idp"]), "order after divide, step %d" % s
        assert np.array_equal(sa["dcell"], sb["dcell"]) and np.array_equal(sa["code"], sb["code"])
        assert np.array_equal(a.begincell(), b.begincell())
        # re-synchronise the float state so that the comparison stays a single-step comparison
        a.upload(sb["idp"], sb["pos"], sb["velrhop"], ib.npb, sb["code"])
        a.symplectic_dtpre = b.symplectic_dtpre
        ia = a.celldiv(False)
        fa, fb = a.forces(), b.forces()
        acea, ara = a.forces_download(); aceb, arb = b.forces_download()
        npb = ib.npb
        tol = tol_for("pb_" if kind == "periodicbox" else "db_")
        assert rl2(acea[npb:], aceb[npb:]) <= tol["ace"], ("ace", s, rl2(acea[npb:], aceb[npb:]))
        if s > 0 or np.linalg.norm(arb) > 1e-3 * np.sqrt(arb.size) * float(cd.cfg.rhopzero):   # v=0 at t=0: Ar is identically ~0 there
            assert rl2(ara, arb) <= tol["ar"], ("ar", s, rl2(ara, arb))
        assert abs(fa.dt - fb.dt) / fb.dt <= tol["dt"]
        dt = b.symplectic_dtpre
        a.step_sympre(dt); b.step_sympre(dt)
        ia, ib = a.celldiv(True), b.celldiv(True)
        sa, sb = a.download(), b.download()
        assert np.array_equal(sa["idp"], sb["idp"]), "order after mid-step divide, step %d" % s
        a.forces(capi.INTERSTEP_SYMCORRECTOR); b.forces(capi.INTERSTEP_SYMCORRECTOR)
        a.step_symcor(dt); b.step_symcor(dt)
        ia, ib = a.celldiv(True), b.celldiv(True)
    a.close(); b.close()


def test_gpu_pips_match_oracle(gpu_lib, oracle_lib):
    cd = dsph.make_case("dambreak", 40000)
    a, b = _both(gpu_lib, oracle_lib, cd)
    b.celldiv(True)
    for _ in range(10):     # leave the initial lattice, where thousands of pairs sit exactly on the 2h cut-off
        b.step()
    sb = b.download()
    a.upload(sb["idp"], sb["pos"], sb["velrhop"], b.npb, sb["code"])
    a.celldiv(False)
    b.forces()
    pa, pb = a.pips(), b.pips()
    assert pa[1] == pb[1] and pa[3] == pb[3], "checked pairs must be identical"
    # accepted pairs may differ only for pairs within rounding of the 2h cut-off (their kernel gradient is ~0)
    assert abs(pa[0] - pb[0]) <= 1e-3 * pb[0] and abs(pa[2] - pb[2]) <= 1e-3 * max(pb[2], 1)


@pytest.mark.parametrize("opts", [dict(), dict(tkernel=1, tvisco=2, visco=0.01, tdensity=1, shiftmode=3, shiftcoef=-2.0), dict(tdensity=3, shiftmode=1, shiftcoef=-2.0),
                                  dict(tkernel=1, tdensity=0)])
def test_gpu_force_variants_agree(gpu_lib, opts):
    """Kernel variants: 1 plain IEEE loop (reference structure and pair order), 4 division-free body in the same loop, 9 deferred pair
    ring with the fp32 candidate test (rows of cells in another order), 10 = production: packed-half candidate culling + hit-mask ring.
    10 accepts exactly the pairs of 9 in the same order with the same arithmetic: BIT-IDENTICAL.  4 and 9 differ from 1 by rounding."""
    cd = dsph.make_case("dambreak", 80000)
    for k, v in opts.items():
        setattr(cd.cfg, k, type(getattr(cd.cfg, k))(v))
    ctx = dsph.create_context(cd)
    assert gpu_lib.lib.dsph_get_force_variant(ctx.h) == 10, "production variant must be the default on this device"
    ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, cd.codes())
    ctx.celldiv(True)
    for _ in range(5):
        ctx.step()
    res = {}
    for variant in (1, 4, 9, 10):
        gpu_lib.lib.dsph_set_force_variant(ctx.h, variant)
        assert gpu_lib.lib.dsph_get_force_variant(ctx.h) == variant
        ctx.celldiv(False)
        fr = ctx.forces()
        res[variant] = (ctx.forces_download(), fr)
    (ace1, ar1), f1 = res[1]
    (ace9, ar9), f9 = res[9]
    (ace10, ar10), f10 = res[10]
    # 10 accepts exactly the pairs of 9, in the same order, with the same source expressions.  In the production instantiation the
    # two kernels also compile to the same roundings (bit-identical); in other instantiations nvcc may contract a*b+c into an FMA in
    # one kernel and not in the other, which is a last-bit difference per pair: 1e-7 is two orders below the parity tolerance.
    d_ace, d_ar = rl2(ace10, ace9), rl2(ar10, ar9)
    print("v10 vs v9:", opts, "ace %.3e ar %.3e" % (d_ace, d_ar), "bit-identical" if np.array_equal(ace10, ace9) and np.array_equal(ar10, ar9) else "")
    if not opts:
        assert np.array_equal(ace10, ace9) and np.array_equal(ar10, ar9), "v10 (packed-half culling) must be bit-identical to v9"
        assert f10.dt == f9.dt and f10.viscdtmax == f9.viscdtmax and f10.acemax == f9.acemax
    assert d_ace < 1e-7 and d_ar < 1e-7 and abs(f10.dt - f9.dt) <= 1e-7 * f9.dt, (d_ace, d_ar)
    for v in (4, 9):
        (ace, ar), f = res[v]
        assert rl2(ace, ace1) < 5e-6 and rl2(ar, ar1) < 5e-6 and abs(f1.dt - f.dt) / f1.dt < 1e-6, (v, rl2(ace, ace1), rl2(ar, ar1))
    ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("opts", [dict(), dict(tvisco=3), dict(tvisco=1, tdensity=1, tkernel=1),
                                  dict(nn_velgrad=2), dict(nn_velgrad=2, tvisco=3, tkernel=1), dict(nn_velgrad=2, tvisco=1, tdensity=2)])
def test_gpu_nn_ring_kernel_matches_plain_loop(gpu_lib, opts):
    """Multiphase FDA and SPH approaches: the production ring kernels (division-free body) against the plain IEEE loops (variant 1)."""
    cd = dsph.make_case("nn2layer", 60000)
    for k, v in opts.items():
        setattr(cd.cfg, k, v)
    ctx = dsph.create_context(cd)
    ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, cd.codes())
    ctx.celldiv(True)
    for _ in range(8):
        ctx.step()
    res = {}
    for variant in (1, 6):
        gpu_lib.lib.dsph_set_force_variant(ctx.h, variant)
        ctx.celldiv(False)
        fr = ctx.forces()
        res[variant] = (ctx.forces_download(), fr)
    (ace1, ar1), f1 = res[1]; (ace6, ar6), f6 = res[6]
    npb = ctx.npb
    assert rl2(ace6[npb:], ace1[npb:]) < 5e-6 and rl2(ar6, ar1) < 5e-6, (rl2(ace6[npb:], ace1[npb:]), rl2(ar6, ar1))
    assert abs(f1.dt - f6.dt) / f1.dt < 1e-5 and abs(f1.viscetadtmax - f6.viscetadtmax) <= 1e-5 * abs(f1.viscetadtmax)


@pytest.mark.parametrize("kind", ["dambreak", "nn2layer"])
def test_gpu_measure_points_match_numpy(gpu_lib, kind):
    """dsph_measure_points (device probes: pressure, velocity, kernel sum) against the numpy interpolation of tests/measure.py."""
    import measure
    cd = dsph.make_case(kind, 60000)
    ctx = dsph.create_context(cd)
    ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, cd.codes())
    ctx.celldiv(True)
    for _ in range(20):
        ctx.step()
    st = ctx.download()
    rng = np.random.default_rng(7)
    lo = np.array(cd.caseposmin); hi = np.array(cd.caseposmax)
    pts = np.vstack([lo + rng.random((60, 3)) * (np.array([0.42, hi[1], 0.32]) - lo),       # inside and around the water column
                     lo + rng.random((10, 3)) * (hi - lo), [[-1.0, 0.1, 0.1], [0.2, 0.3, 5.0]]])  # anywhere in the tank, outside the domain
    press, vel, kc = ctx.measure_points(pts)
    fl = (st["code"] & capi.CODE_MASKTYPE) == capi.CODE_TYPE_FLUID
    pf = st["pos"][fl]; rf = st["velrhop"][fl, 3]; vf = st["velrhop"][fl, :3]
    cfg = cd.cfg
    if kind == "nn2layer":        # per-phase EOS and mass
        ph = (st["code"][fl] & 0x7ff).astype(int)
        rho0 = np.array([cfg.phases[k].rho for k in range(cfg.nn_nphases)])[ph]
        cs0 = np.array([cfg.phases[k].cs0 for k in range(cfg.nn_nphases)])[ph]
        gam = np.array([cfg.phases[k].gamma for k in range(cfg.nn_nphases)])[ph]
        pj = (cs0 ** 2 * rho0 / gam) * ((rf.astype(np.float64) / rho0) ** gam - 1.0)
        mass = rho0 * cfg.dp ** 3
    else:
        pj = float(cfg.cteb) * ((rf.astype(np.float64) / float(cfg.rhopzero)) ** float(cfg.gamma) - 1.0)
        mass = np.full(len(rf), float(cfg.massfluid))
    h = float(cfg.kernelh)
    refp = np.zeros(len(pts)); refv = np.zeros((len(pts), 3)); refk = np.zeros(len(pts))
    for k, pt in enumerate(pts):
        d = pf - pt; near = np.all(np.abs(d) <= 2 * h, axis=1)
        if near.any():
            r = np.sqrt((d[near] ** 2).sum(axis=1)); w = measure.wendland(r / h, h) * (mass[near] / rf[near])
            refk[k] = w.sum()
            if refk[k] > 0:
                refp[k] = (w * pj[near]).sum() / refk[k]; refv[k] = (w[:, None] * vf[near]).sum(axis=0) / refk[k]
    assert np.count_nonzero(refk > 0.5) > 30, "most probes must lie inside the fluid"
    assert kc[-1] == 0 and kc[-2] == 0 and press[-1] == 0
    pscale = max(np.abs(refp).max(), 1.0); vscale = max(np.abs(refv).max(), 1e-6)
    assert np.max(np.abs(kc - refk)) <= 2e-5 * max(refk.max(), 1.0), np.max(np.abs(kc - refk))
    assert np.max(np.abs(press - refp)) <= 1e-4 * pscale, (np.max(np.abs(press - refp)), pscale)
    assert np.max(np.abs(vel - refv)) <= 1e-4 * vscale, (np.max(np.abs(vel - refv)), vscale)
    ctx.close()


def test_gpu_large_case_properties(gpu_lib):
    """Full-size properties (no CPU oracle at this size): sortedness, permutation, begin-cell consistency, idempotence."""
    cd = dsph.make_case("dambreak", 2_000_000)
    ctx = dsph.create_context(cd)
    ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, cd.codes())
    info = ctx.celldiv(True)
    st = ctx.download()
    assert info.np == cd.np_
    assert np.array_equal(np.sort(st["idp"]), np.arange(cd.np_, dtype=np.uint32)), "reorder must be a permutation"
    beg = ctx.begincell()
    assert np.all(np.diff(beg.astype(np.int64)) >= 0) and beg[0] == 0 and beg[2 * info.nct + 1] == info.np
    # every particle lies in the box its cell says
    k = ctx  # noqa: F841
    dcc = int(info.domcellcode)
    cel = st["dcell"]
    cx = cel >> ((dcc >> 10) & 31); cy = ((cel << np.uint32(dcc >> 25)) & 0xFFFFFFFF) >> ((dcc >> 5) & 31)
    cz = ((cel << np.uint32(dcc & 31)) & 0xFFFFFFFF) >> (dcc & 31)
    cx = cx.astype(np.int64); cy = cy.astype(np.int64); cz = cz.astype(np.int64)
    box = (cx - info.cellmin[0]) + (cy - info.cellmin[1]) * info.ncx + (cz - info.cellmin[2]) * info.ncx * info.ncy
    inside = ((cx >= info.cellmin[0]) & (cx <= info.cellmax[0]) & (cy >= info.cellmin[1]) & (cy <= info.cellmax[1])
              & (cz >= info.cellmin[2]) & (cz <= info.cellmax[2]))
    fluid = (st["code"] & capi.CODE_MASKTYPE) == capi.CODE_TYPE_FLUID
    assert np.all(inside[fluid])
    key = np.where(fluid, box + info.nct + 1, np.where(inside, box, info.nct))   # boundary outside the cell domain: BoxBoundIgnore
    assert np.all(np.diff(key.astype(np.int64)) >= 0), "particles must be sorted by box"
    # stable: ids inside one box keep ascending order on the first divide (input was id-ordered)
    same = np.diff(key.astype(np.int64)) == 0
    assert np.all(np.diff(st["idp"].astype(np.int64))[same] > 0), "sort must be stable"
    # idempotence: dividing a sorted state changes nothing
    ctx.celldiv(True)
    st2 = ctx.download()
    assert np.array_equal(st2["idp"], st["idp"]) and np.array_equal(ctx.begincell(), beg)
    fr = ctx.forces()
    ace, ar = ctx.forces_download()
    assert np.all(np.isfinite(ace)) and np.all(np.isfinite(ar)) and fr.dt > 0
    ctx.close()


@pytest.mark.parametrize("kind", ["dambreak", "nn2layer", "dambreak:tstep=1,verletsteps=5,tvisco=2,visco=0.01,shiftmode=3,shiftcoef=-2",
                                  "dambreak@rebalance", "nn2layer@rebalance", "nn2layer:nn_velgrad=2", "nn2layer:nn_velgrad=2,tvisco=3",
                                  # periodic along y and z + slabs along x (periodic copies of owned particles, migrants and halo copies alike)
                                  "periodicbox:perix=0", "periodicbox:perix=0,tstep=1,verletsteps=5",
                                  # periodic along x too: the two ranks form a ring, halo copies cross the periodic face shifted by the period
                                  "periodicbox", "periodicbox@rebalance", "periodicbox:tstep=1,verletsteps=5"])
def test_gpu_two_slabs_equal_single_gpu(kind):
    """x-slab decomposition over 2 GPUs (NCCL send/recv halos + all-reduce dt) reproduces the single-GPU run."""
    import subprocess, sys, os, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", os.path.join(root, "tests", "mgpu_gpu_check.py"), "200000", "10", kind],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("mode", ["peer_memory", "nccl_sendrecv", "fallback"])
def test_gpu_two_slabs_record_transports(mode):
    """The three ways the halo / migrant records travel give the single-GPU result: written by the pack kernel into the neighbour's
    buffer (peer memory, default), ncclSend/ncclRecv (DSPH_MGPU_DIRECT=0), and the per-exchange fall-back when the records do not
    fit the peer buffer (forced with a 64-record buffer)."""
    import subprocess, sys, os, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
    env = dict(os.environ, DSPH_MGPU_EXPECT=mode)
    if mode == "nccl_sendrecv":
        env["DSPH_MGPU_DIRECT"] = "0"
    if mode == "fallback":
        env["DSPH_MGPU_DIRECT_CAP"] = "64"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tests", "mgpu_gpu_check.py"), "200000", "10", "dambreak@rebalance"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, env=env)
    assert "MGPU_CHECK PASS" in r.stdout, r.stdout[-3000:]


def _matrix():
    """Option combinations across kernel x viscosity x DDT x shifting x step algorithm x cell mode (+ multiphase variants)."""
    combos = []
    i = 0
    for tkernel in (1, 2):
        for tvisco in (1, 2):
            for tdensity in (0, 1, 2, 3):
                shift = (0, 1, 2, 3)[i % 4]; tstep = (1, 2)[(i // 2) % 2]; cellmode = 2 if i % 5 == 4 else 1
                combos.append(("dambreak", dict(tkernel=tkernel, tvisco=tvisco, visco=0.1 if tvisco == 1 else 0.01, tdensity=tdensity,
                                                shiftmode=shift, shiftcoef=-2.0, shifttfs=(1.5 if i % 3 == 0 else 0.0), tstep=tstep,
                                                verletsteps=4, cellmode=cellmode)))
                i += 1
    for velgrad in (1, 2):
        for tvisco in (1, 2, 3):
            for tkernel in (1, 2):
                combos.append(("nn2layer", dict(nn_velgrad=velgrad, tvisco=tvisco, tkernel=tkernel, tdensity=(1, 2, 3, 0)[i % 4],
                                                shiftmode=(3, 0, 1)[i % 3], tstep=(2, 1)[i % 2], verletsteps=4)))
                i += 1
    combos.append(("dambreak2d", dict(tkernel=2, tvisco=1, tdensity=2, shiftmode=3, shiftcoef=-2.0)))
    combos.append(("nn2layer2d", dict(nn_velgrad=1, tvisco=2, tdensity=3)))
    return combos


@pytest.mark.parametrize("kind,opts", _matrix(), ids=lambda v: v if isinstance(v, str) else "-".join("%s%g" % (k[:4], x) for k, x in v.items()))
def test_gpu_option_matrix_vs_oracle(gpu_lib, oracle_lib, kind, opts):
    """Every template path of the production kernels against the CPU checker on a small case: a few whole steps on both sides,
    order bit-exact after every divide, then one force evaluation from the same state within the single-step tolerance."""
    cd = dsph.make_case(kind, 6000)
    for k, v in opts.items():
        setattr(cd.cfg, k, type(getattr(cd.cfg, k))(v))
    a, b = _both(gpu_lib, oracle_lib, cd)
    a.celldiv(True); b.celldiv(True)
    for s in range(5):
        a.step(); b.step()
        sa, sb = a.download(), b.download()
        assert a.np_ == b.np_ and np.array_equal(sa["idp"], sb["idp"]), "order after step %d" % s
    # same state on both sides, one evaluation
    a.upload(sb["idp"], sb["pos"], sb["velrhop"], b.npb, sb["code"])
    if cd.cfg.tstep == capi.STEP_SYMPLECTIC:
        a.symplectic_dtpre = b.symplectic_dtpre
    b2 = SphContext(oracle_lib, cd.cfg)
    b2.upload(sb["idp"], sb["pos"], sb["velrhop"], b.npb, sb["code"])
    a.celldiv(False); b2.celldiv(False)
    fa, fb = a.forces(), b2.forces()
    acea, ara = a.forces_download(); aceb, arb = b2.forces_download()
    npb = b2.npb
    assert rl2(acea[npb:], aceb[npb:]) <= 1e-5 and rl2(ara, arb) <= 1e-5, (rl2(acea[npb:], aceb[npb:]), rl2(ara, arb))
    assert abs(fa.dt - fb.dt) <= 1e-5 * fb.dt
    a.close(); b.close(); b2.close()


@pytest.mark.parametrize("tstep", [capi.STEP_SYMPLECTIC, capi.STEP_VERLET])
def test_gpu_excluded_particles_match_oracle(gpu_lib, oracle_lib, tstep):
    """Particles that leave: displacement above MovLimit (OUTMOVE), density outside RhopOut (OUTRHOP), position outside the map
    (OUTPOS).  The divide must drop the same particles on both sides (S/JSphCpu.cpp:1240-1293, S/JCellDivCpuSingle.cpp:134-194)."""
    cd = dsph.make_case("dambreak", 6000)
    cd.cfg.tstep = tstep
    a, b = _both(gpu_lib, oracle_lib, cd)
    a.celldiv(True); b.celldiv(True)
    for _ in range(2):
        a.step(); b.step()
    st = b.download()
    fl = np.nonzero((st["code"] & capi.CODE_MASKTYPE) == capi.CODE_TYPE_FLUID)[0]
    vr = st["velrhop"].copy(); pos = st["pos"].copy()
    vr[fl[10], 0] = 5.0e4                                  # moves more than 0.9*Scell in one step
    vr[fl[200], 3] = 5000.0                                # density out of [700,1300]
    zmax = float(cd.cfg.maprealposmax[2])
    pos[fl[400], 2] = zmax - 1e-7; vr[fl[400], 2] = 80.0   # crosses the top of the map
    for c in (a, b):
        c.upload(st["idp"], pos, vr, b.npb, st["code"])
        c.celldiv(True)
    n0 = b.np_
    for s in range(3):
        a.step(); b.step()
        sa, sb = a.download(), b.download()
        assert a.np_ == b.np_, ("np after step %d" % s, a.np_, b.np_)
        assert np.array_equal(sa["idp"], sb["idp"]) and np.array_equal(sa["code"], sb["code"])
    assert n0 - b.np_ >= 3, (n0, b.np_)          # the three tagged particles (and whatever the density spike blew away) are gone
    gone = set(st["idp"][[fl[10], fl[200], fl[400]]].tolist())
    assert gone.isdisjoint(set(sb["idp"].tolist()))
    a.close(); b.close()


# ---- the reference solver run LIVE on the GPU box at the configuration sizes ---------------------------------
@pytest.mark.parametrize("name,np_target,steps,nsteps", [("C1_dambreak_170k", 170000, [0, 12], 13), ("DB_2M", 2000000, [3], 4)])
def test_gpu_stages_vs_live_reference(gpu_lib, name, np_target, steps, nsteps):
    """C1 (BASELINE.json configs[0]: ~170 k particles, Symplectic + Wendland + delta-SPH) and DB(2 M): every stage of the dumped
    steps -- divide (bit-exact order, cells, begin-cell table), predictor/corrector forces (Ace/Ar/dt <= 1e-5), updates -- against
    the reference's JSphCpuSingle executed now, on this box."""
    import live_reference
    if not live_reference.available():
        pytest.skip("oracle/_ref/ref_harness52 or bin/mkcase missing")
    g, cd = live_reference.live_golden(name, "dambreak", np_target, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2"], steps, nsteps)
    print(name, "np", cd.np_, "boundary fraction %.3f" % (cd.npb / cd.np_))
    for step in g.steps:
        res = parity.run_step_stages(gpu_lib, g, step, TOL, capacity=int(cd.np_ * 1.1) + 1000)
        print(name, step, res)
    if nsteps <= 20 and cd.np_ < 500000:
        out = parity.run_trajectory(gpu_lib, g, capacity=int(cd.np_ * 1.1) + 1000)
        print(name, "trajectory", out)
        assert out["np"][0] == out["np"][1] and out["time"] < 1e-5 and out["pos"] < 1e-6 and out["velrhop"] < 1e-4


def test_gpu_moving_piston_vs_live_reference(gpu_lib):
    """Moving boundaries at 150 k particles against the reference's JSphCpuSingle executed now: an oscillating piston next to the water
    column (mvrectsinu).  Stages of two steps (the motion stage bit-exact) and the 15-step trajectory with the library evaluating the
    motion laws of Case.xml itself."""
    import live_reference
    if not live_reference.available():
        pytest.skip("oracle/_ref/ref_harness52 or bin/mkcase missing")
    g, cd = live_reference.live_golden("movpiston_150k", "movpiston", 150000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2"], [0, 14], 15)
    assert int(cd.cfg.casenmoving) > 1000
    for step in g.steps:
        # VelMax of the chained corrector stage follows the predictor's velocities (dt * Ace at step 0), i.e. the Ace tolerance
        res = parity.run_step_stages(gpu_lib, g, step, dict(TOL, velmax=5e-6), capacity=int(cd.np_ * 1.1) + 1000)
        print("movpiston_150k", step, res)
    out = parity.run_trajectory(gpu_lib, g, capacity=int(cd.np_ * 1.1) + 1000)
    print("movpiston_150k trajectory", out)
    assert out["np"][0] == out["np"][1] and out["time"] < 1e-5 and out["pos"] < 1e-6 and out["velrhop"] < 1e-4


def test_gpu_production_kernel_vs_ieee_kernel_at_20m(gpu_lib):
    """DB(20 M), the benchmark state: the production kernel (variant 10: packed-half culling, deferred pair ring,
    division-free body, row order of its own) against variant 1 (plain IEEE loop in the reference's structure and pair order, the
    kernel the 17 reference dumps pin) on the same divide: Ace/Ar/dt within 1e-5."""
    cd = dsph.make_case("dambreak", 20_000_000)
    ctx = dsph.create_context(cd)
    ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, cd.codes())
    ctx.celldiv(True)
    for _ in range(3):
        ctx.step()
    res = {}
    for variant in (1, 10):
        gpu_lib.lib.dsph_set_force_variant(ctx.h, variant)
        ctx.celldiv(False)
        fr = ctx.forces()
        res[variant] = (ctx.forces_download(), fr)
    (ace1, ar1), f1 = res[1]
    (ace10, ar10), f10 = res[10]
    npb = ctx.npb
    out = dict(ace=rl2(ace10[npb:], ace1[npb:]), ar=rl2(ar10, ar1), dt=abs(f10.dt - f1.dt) / f1.dt, np=ctx.np_)
    print("DB(20 M) production vs IEEE kernel:", out)
    ctx.close()
    assert out["ace"] <= TOL["ace"] and out["ar"] <= TOL["ar"] and out["dt"] <= TOL["dt"], out
