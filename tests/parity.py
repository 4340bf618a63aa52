"""Stage-by-stage parity of one solver (CPU checker or CUDA path) against a golden dump of the reference solver.

Integer/index work (cell assignment, sorted order, begin-cell table, particle codes) must be bit-exact; floating
point results are compared with a relative L2 norm (north_star: 1e-5 for single-step Ace/Ar/dt in fp32)."""
import ctypes as C
import os

import numpy as np

from dualsphysics_multilayer_b200 import capi
from dualsphysics_multilayer_b200.capi import Config, SphContext

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
GOLDEN_CASES = ["db_sym_wend_art_ddt2", "db_ver_cubic_art_ddt1_shift", "db_sym_wend_lamsps_ddt3",
                "db_sym_cubic_art_noddt_half", "pb_sym_wend_art_ddt1", "pb_ver_wend_art_ddt1",
                "nn_sym_fda_lam_ddt3_shift", "nn_ver_sph_lam_ddt1", "nn_sym_sph_conseq_ddt2_shift", "nn_sym_fda_conseq_cubic",
                "nn_sym_fda_art_ddt1", "db2d_sym_wend_art_ddt2_shift", "db2d_ver_cubic_lamsps_ddt1", "nn2d_sym_sph_lam_ddt3_shift",
                "nn2d_sym_fda_conseq_ddt2", "db_sym_wend_art_ddt2_rhopout", "db_ver_wend_art_ddt1_rhopout",
                "nn3_2d_sym_fda_lam_ddt3_shift", "nn3_sym_sph_conseq_ddt2"]
# mDBC boundary correction: dumps of the reference with Boundary=2; the CPU checker has no mDBC, the CUDA path is compared with the
# reference directly (tests/test_gpu_parity.py)
# moving boundaries: the mDBC one needs normals, which the CPU checker does not carry
MOVING_CASES = ["mv_sym_gate_rect_wait", "mv_ver_flap_rotsinu", "mv_sym_piston_rectsinu_cubic"]
MOVING_MDBC_CASES = ["mv2d_sym_flap_rotsinu_mdbc"]
MDBC_CASES = ["db_sym_wend_art_ddt1_mdbc", "db_sym_cubic_art_ddt2_mdbccorr", "db2d_ver_wend_art_ddt1_mdbc"]
GPU_PENDING = []                             # cases the CUDA path does not cover yet


PROPAGATION_TOL = 1e-4      # forces evaluated on a chained (not re-synchronised) state, see run_step_stages


def rl2(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


class _Pack(dict):
    """In-memory stand-in for a loaded npz (live dumps of the reference harness, tests/live_reference.py)."""
    @property
    def files(self):
        return list(self.keys())


class Golden:
    def __init__(self, name, pack=None):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz")) if pack is None else _Pack(pack)
        self.cfg = Config.from_buffer_copy(self.z["config"].tobytes())
        self.steps = [int(s) for s in self.z["steps"]]
        self.nsteps = int(self.z["nsteps"])
        self.symplectic = (self.cfg.tstep == capi.STEP_SYMPLECTIC)
        self.periodic = bool(self.cfg.perix or self.cfg.periy or self.cfg.periz)

    def stage(self, step, stage):
        pre = "s%04d_%s/" % (step, stage)
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def has(self, step, stage):
        pre = "s%04d_%s/" % (step, stage)
        return any(k.startswith(pre) for k in self.z.files)


def make_ctx(lib, cfg, capacity=0):
    """Context for a comparison with dumps of the reference BINARY: the CUDA library is asked to reproduce that binary's
    index-dependent pressure rounding (dsph_set_refbinary_tail, off in production); the CPU checker does it by default."""
    ctx = SphContext(lib, cfg, capacity)
    if lib.prefix == "dsph_":
        fn = lib.lib.dsph_set_refbinary_tail
        fn.argtypes = [C.c_void_p, C.c_int]
        fn(ctx.h, 1)
    return ctx


def upload_state(ctx, st, cfg):
    """Uploads a dumped (cell-sorted) reference state, periodic duplicates included."""
    code = st["code"]
    keep = np.ones(code.shape[0], bool)          # periodic duplicates are kept: the first divide runs with updateperiodic=0
    npb = int(st["npb"])
    ctx.upload(st["idp"][keep], st["pos"][keep], st["velrhop"][keep], npb, code[keep])
    if "velrhopm1" in st:
        ctx.set_velrhopm1(st["velrhopm1"][keep])
    if "spstau" in st:
        ctx.set_spstau(st["spstau"][keep])
    if "boundnormal" in st:                       # mDBC: vectors particle -> ghost node in the dumped particle order
        ctx.set_boundnormals(st["boundnormal"][keep][:npb])
    if cfg.tstep == capi.STEP_SYMPLECTIC:
        ctx.symplectic_dtpre = float(st["symplecticdtpre"])
    else:
        fn = ctx.lib.fn("set_verletstep")
        fn.argtypes = [C.c_void_p, C.c_int]
        fn(ctx.h, int(st["verletstep"]))
    return keep


def check_divide(ctx, ref, exact_ints=True):
    """After ctx.celldiv(): cell domain, counts, sorted order, dcell, codes and begin-cell table must equal the dump."""
    info = ctx.divinfo
    res = {}
    assert int(info.np) == int(ref["np"]), ("np", info.np, ref["np"])
    assert int(info.npb) == int(ref["npb"]) and int(info.npbok) == int(ref["npbok"])
    assert list(info.cellmin) + list(info.cellmax) == [int(v) for v in ref["celldomain"]], "cell domain"
    assert [info.ncx, info.ncy, info.ncz, info.nct] == [int(v) for v in ref["ncells"]]
    assert int(info.domcellcode) == int(ref["domcellcode"])
    st = ctx.download()
    assert np.array_equal(st["idp"], ref["idp"]), "sorted particle order (idp)"
    assert np.array_equal(st["dcell"], ref["dcell"]), "dcell"
    assert np.array_equal(st["code"], ref["code"]), "code"
    # boxes of valid particles: nct boundary cells, BoxBoundIgnore, nct fluid cells, begin of BoxBoundOut (= np).  The
    # out/ignore boxes after them hold particles the reference dropped in this divide, which a replay does not have.
    nvalid = 2 * int(info.nct) + 2
    assert np.array_equal(ctx.begincell()[:nvalid], ref["begincell"][:nvalid]), "begin-cell table"
    assert np.array_equal(st["pos"], ref["pos"]), "positions must be moved bit-exactly by the reorder"
    assert np.array_equal(st["velrhop"], ref["velrhop"])
    return res


def check_forces(ctx, ref, cfg, tol, interstep):
    fr = ctx.forces(interstep)
    ace, ar = ctx.forces_download()
    npb = int(ref["npb"])
    out = dict(ace=rl2(ace[npb:], ref["ace"][npb:]), ar=rl2(ar, ref["ar"]),
               dt=abs(fr.dt - float(ref["dt"])) / float(ref["dt"]),
               viscdt=abs(fr.viscdtmax - float(ref["viscdtmax"])) / max(abs(float(ref["viscdtmax"])), 1e-30),
               acemax=abs(fr.acemax - float(ref["acemax"])) / max(float(ref["acemax"]), 1e-30),
               velmax=abs(fr.velmax - float(ref["velmax"])) / max(float(ref["velmax"]), 1e-30))
    if "velrhop" in ref:                          # mDBC: densities of the boundary particles after the correction
        st = ctx.download()
        out["mdbc_rhop"] = rl2(st["velrhop"][:npb, 3] - float(cfg.rhopzero), ref["velrhop"][:npb, 3] - float(cfg.rhopzero))
    for k, v in out.items():
        assert v <= tol.get(k, tol["default"]), (k, v, out)
    return fr, out


def run_step_stages(lib, g: Golden, step, tol, capacity=0):
    """Replays one dumped time step stage by stage, re-synchronising with the reference state before each stage."""
    cfg = g.cfg
    results = {}
    ctx = make_ctx(lib, cfg, capacity)
    s0 = g.stage(step, "div0")
    upload_state(ctx, s0, cfg)
    ctx.celldiv(False)                         # the state already holds its periodic duplicates
    check_divide(ctx, s0)                      # divide of the reference's own state: must reproduce it exactly
    t2 = tol                                   # one tolerance for every step (round 1 did not check Ar at step 0)
    if g.symplectic:
        dt = float(s0["symplecticdtpre"])
        _, results["fpre"] = check_forces(ctx, g.stage(step, "fpre"), cfg, t2, capi.INTERSTEP_SYMPREDICTOR)
        ctx.step_sympre(dt)
        u = g.stage(step, "upre")
        st = ctx.download()
        normal = (st["code"] & capi.CODE_MASKSPECIAL) != capi.CODE_PERIODIC
        assert np.array_equal(st["dcell"][normal], u["dcell"][normal]), "dcell after predictor"
        assert np.array_equal(st["code"], u["code"]), "code after predictor"
        results["upre_pos"] = float(np.abs(st["pos"][normal] - u["pos"][normal]).max())
        results["upre_velrhop"] = rl2(st["velrhop"], u["velrhop"])
        assert results["upre_pos"] <= tol["pos"] and results["upre_velrhop"] <= tol["state"], results
        ctx.celldiv(True)
        d1 = g.stage(step, "div1")
        st = ctx.download()
        assert np.array_equal(st["idp"], d1["idp"]) and np.array_equal(ctx.begincell(), d1["begincell"]), "mid-step divide"
        # Corrector forces, the parity criterion: evaluated on the REFERENCE's own mid-step state (same inputs -> outputs within tol).
        c2 = make_ctx(lib, cfg, capacity)
        upload_state(c2, d1, cfg)
        c2.celldiv(False)
        _, results["fcor"] = check_forces(c2, g.stage(step, "fcor"), cfg, tol, capi.INTERSTEP_SYMCORRECTOR)
        c2.close()
        # The chained context (its mid-step state is this solver's own predictor result, equal to the reference's to 1e-6) goes on to
        # feed the corrector update.  Its forces are bounded too, but by a propagation bound, not by the parity tolerance: Tait's
        # equation turns a last-bit difference of rho into a relative pressure difference of 2^-23/(rho/rho0-1), which in a nearly
        # uniform field (periodic box at step 0: rho/rho0-1 ~ 1e-4) is 1e-3 of the pressure.
        _, results["fcor_chained"] = check_forces(ctx, g.stage(step, "fcor"), cfg, dict(tol, ace=PROPAGATION_TOL, ar=PROPAGATION_TOL), capi.INTERSTEP_SYMCORRECTOR)
        ctx.step_symcor(dt)
        u = g.stage(step, "ucor")
    else:
        fr, results["fver"] = check_forces(ctx, g.stage(step, "fver"), cfg, t2, capi.INTERSTEP_VERLET)
        ctx.step_verlet(float(g.stage(step, "fver")["dt"]))
        u = g.stage(step, "uver")
    st = ctx.download()
    normal = (st["code"] & capi.CODE_MASKSPECIAL) != capi.CODE_PERIODIC
    assert np.array_equal(st["code"], u["code"]), "code after update"
    results["u_pos"] = float(np.abs(st["pos"][normal] - u["pos"][normal]).max())
    results["u_velrhop"] = rl2(st["velrhop"], u["velrhop"])
    # the new cell of a particle may differ only if its position differs (cell boundary crossing by rounding)
    same = np.all(st["pos"] == u["pos"], axis=1) & normal
    assert np.array_equal(st["dcell"][same], u["dcell"][same]), "dcell after update"
    assert results["u_pos"] <= tol["pos"] and results["u_velrhop"] <= tol["state"], results
    ctx.close()
    if g.has(step, "motion") and g.has(step, "dmot"):
        # RunMotion + RunCellDivide of the reference's own updated state with the motion data JDsMotion computed: moved positions,
        # velocities, cells and the re-sorted boundary must be those of the dump, bit for bit (double adds, unfused matrix products)
        stepdt = float(g.stage(step, "end")["stepdt"])
        c3 = make_ctx(lib, cfg, capacity)
        upload_state(c3, u, cfg)
        c3.motion_run(motion_from_dump(g.stage(step, "motion")), stepdt)
        c3.celldiv(True)
        dm = g.stage(step, "dmot")
        check_divide(c3, dm)
        if "boundnormal" in dm:
            npb = int(dm["npb"])
            results["motion_normals"] = float(np.abs(c3.boundnormals(npb) - dm["boundnormal"][:npb]).max())
            assert results["motion_normals"] <= 1e-7, results
        c3.close()
    return results


def motion_from_dump(m):
    """s<step>_motion dump of the harness (StMotionData of every object) -> list of capi.MotionData."""
    nobj = int(m["nobj"])
    head = np.asarray(m["head"]).reshape(nobj, 3); val = np.asarray(m["val"]).reshape(nobj, 22)
    out = []
    for i in range(nobj):
        d = capi.MotionData()
        d.idbegin, d.count, d.type = int(head[i, 0]), int(head[i, 1]), int(head[i, 2])
        d.linmov[:] = list(val[i, 0:3]); d.linvel[:] = list(val[i, 3:6]); d.matmov[:] = list(val[i, 6:22])
        out.append(d)
    return out


def motion_log(g):
    """Per-step log of the harness: (timestep, dt, [type, linmov3, linvel3, matmov16] per object) for every step of the run."""
    z = g.z
    key = "s%04d_motionlog" % g.nsteps
    nobj = int(z[key + "/nobj"])
    tdt = np.asarray(z[key + "/t_dt"]).reshape(-1, 2)
    val = np.asarray(z[key + "/val"]).reshape(tdt.shape[0], nobj, 23)
    return tdt, val


def write_case_xml(g, directory, casename="Case"):
    """Case.xml of a golden (stored for the cases whose host-side definitions live in it: gauges, motion)."""
    import os
    with open(os.path.join(directory, casename + ".xml"), "wb") as f:
        f.write(bytes(g.z["case/xml"]))


def run_trajectory(lib, g: Golden, capacity=0):
    """Whole-step loop from the initial case for g.nsteps steps; returns drift vs the reference's final state."""
    cfg = g.cfg
    ctx = make_ctx(lib, cfg, capacity)
    z = g.z
    idp = z["case/idp"]; npb = int(z["case/npb"])
    if "case/code" in z.files:               # multiphase cases: the phase is the type value of the code
        code = z["case/code"]
    else:
        code = np.where(idp < cfg.casenfixed, capi.CODE_TYPE_FIXED, capi.CODE_TYPE_FLUID).astype(np.uint16)
    ctx.upload(idp, z["case/pos"], z["case/velrhop"], npb, code)
    if "case/normals" in z.files:                 # mDBC: file vectors (particle -> boundary limit) doubled, JSph::ConfigBoundNormals
        ctx.set_boundnormals((2.0 * z["case/normals"][idp[:npb]]).astype(np.float32))
    ctx.celldiv(True)
    t = 0.0
    moving = int(cfg.casenmoving) > 0
    if moving and lib.prefix == "dsph_":
        # the product evaluates the motion laws of Case.xml itself (host/motion.cpp) inside dsph_step
        import tempfile
        import dualsphysics_multilayer_b200 as dsph
        with tempfile.TemporaryDirectory() as tmp:
            write_case_xml(g, tmp)
            motion = dsph.CaseMotion(tmp, "Case")
        ctx.motion_attach(motion)
    if moving and lib.prefix != "dsph_":
        # the checker is driven stage by stage with the motion data the reference logged for every step
        tdt, val = motion_log(g)
        head = np.asarray(g.stage(int(g.steps[0]), "motion")["head"]).reshape(-1, 3)
        for s in range(g.nsteps):
            objs = []
            for i in range(val.shape[1]):
                d = capi.MotionData(); d.idbegin, d.count, d.type = int(head[i, 0]), int(head[i, 1]), int(val[s, i, 0])
                d.linmov[:] = list(val[s, i, 1:4]); d.linvel[:] = list(val[s, i, 4:7]); d.matmov[:] = list(val[s, i, 7:23])
                objs.append(d)
            if g.symplectic:
                dt = ctx.symplectic_dtpre
                fp = ctx.forces(capi.INTERSTEP_SYMPREDICTOR); ctx.step_sympre(dt); ctx.celldiv(True)
                fc = ctx.forces(capi.INTERSTEP_SYMCORRECTOR); ctx.step_symcor(dt)
                ctx.symplectic_dtpre = min(fp.dt, fc.dt)
            else:
                dt = ctx.forces(capi.INTERSTEP_VERLET).dt
                ctx.step_verlet(dt)
            ctx.motion_run(objs, dt)
            ctx.celldiv(True)
            t += dt
    else:
        for _ in range(g.nsteps):
            dt, _ = ctx.step()
            t += dt
    fin = g.stage(g.nsteps, "final")
    st = ctx.download()
    out = dict(time=abs(t - float(fin["timestep"])) / float(fin["timestep"]), np=(ctx.np_, int(fin["np"])))
    # compare per particle id (order may differ if a particle crossed a cell boundary by rounding)
    o1 = np.argsort(st["idp"], kind="stable"); o2 = np.argsort(fin["idp"], kind="stable")
    if ctx.np_ == int(fin["np"]) and np.array_equal(st["idp"][o1], fin["idp"][o2]):
        normal = (fin["code"][o2] & capi.CODE_MASKSPECIAL) == capi.CODE_NORMAL
        out["pos"] = float(np.abs(st["pos"][o1][normal] - fin["pos"][o2][normal]).max())
        out["velrhop"] = rl2(st["velrhop"][o1][normal], fin["velrhop"][o2][normal])
        out["order_equal"] = bool(np.array_equal(st["idp"], fin["idp"]))
    ctx.close()
    return out
