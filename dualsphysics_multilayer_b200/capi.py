"""ctypes view of include/dsph_b200.h.

The structures below mirror the C PODs field by field.  `SphLib` wraps one shared library that exports the
hot-path entry points with a given symbol prefix: the product library ``libdsph_b200.so`` uses ``dsph_``.
(tests drive the CPU checker in ``oracle/`` through the same class with prefix ``orc_`` -- the product never
loads it.)  There is no CPU fallback: creating a context without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

MAXPHASES = 10

# typecode bits (S/DualSphDef.h:164-226)
CODE_MASKSPECIAL = 0xE000
CODE_NORMAL = 0x0000
CODE_PERIODIC = 0x2000
CODE_OUTIGNORE = 0x4000
CODE_OUTMOVE = 0x6000
CODE_OUTPOS = 0x8000
CODE_OUTRHOP = 0xA000
CODE_MASKTYPE = 0x1800
CODE_TYPE_FIXED = 0x0000
CODE_TYPE_MOVING = 0x0800
CODE_TYPE_FLUID = 0x1800
CODE_MASKVALUE = 0x07FF

KERNEL_CUBIC, KERNEL_WENDLAND = 1, 2
STEP_VERLET, STEP_SYMPLECTIC = 1, 2
VISCO_ARTIFICIAL, VISCO_LAMINARSPS = 1, 2
INTERSTEP_VERLET, INTERSTEP_SYMPREDICTOR, INTERSTEP_SYMCORRECTOR = 1, 2, 3


class Phase(C.Structure):
    _fields_ = [("mkfluid", C.c_uint32), ("rho", C.c_float), ("mass", C.c_float), ("cs0", C.c_float),
                ("cteb", C.c_float), ("gamma", C.c_float), ("visco", C.c_float), ("tau_yield", C.c_float),
                ("tau_max", C.c_float), ("bi_multi", C.c_float), ("m_nn", C.c_float), ("n_nn", C.c_float),
                ("phasetype", C.c_uint32), ("reserved", C.c_uint32)]


class Config(C.Structure):
    _fields_ = [
        ("tkernel", C.c_int32), ("tstep", C.c_int32), ("tvisco", C.c_int32), ("tdensity", C.c_int32),
        ("shiftmode", C.c_int32), ("cellmode", C.c_int32), ("verletsteps", C.c_int32), ("simulate2d", C.c_int32),
        ("perix", C.c_int32), ("periy", C.c_int32), ("periz", C.c_int32),
        ("dtallparticles", C.c_int32), ("dtfixed_on", C.c_int32), ("stable", C.c_int32),
        ("nn_on", C.c_int32), ("nn_velgrad", C.c_int32), ("nn_nphases", C.c_int32), ("tboundary", C.c_int32),
        ("kernelh", C.c_float), ("cteb", C.c_float), ("gamma", C.c_float), ("rhopzero", C.c_float),
        ("massfluid", C.c_float), ("massbound", C.c_float),
        ("gravity", C.c_float * 3),
        ("visco", C.c_float), ("viscoboundfactor", C.c_float), ("ddtvalue", C.c_float),
        ("shiftcoef", C.c_float), ("shifttfs", C.c_float),
        ("coefdtmin", C.c_float), ("rhopoutmin", C.c_float), ("rhopoutmax", C.c_float), ("nn_relaxationdt", C.c_float),
        ("dp", C.c_double), ("cflnumber", C.c_double), ("dtini", C.c_double), ("dtmin", C.c_double),
        ("dtfixed", C.c_double), ("simulate2dposy", C.c_double),
        ("timemax", C.c_double), ("timeout", C.c_double),
        ("maprealposmin", C.c_double * 3), ("maprealposmax", C.c_double * 3),
        ("perixinc", C.c_double * 3), ("periyinc", C.c_double * 3), ("perizinc", C.c_double * 3),
        ("casenp", C.c_uint32), ("casenfixed", C.c_uint32), ("casenmoving", C.c_uint32),
        ("casenfloat", C.c_uint32), ("casenfluid", C.c_uint32),
        ("partsoutmax_on", C.c_uint32), ("partsoutmax", C.c_float),
        ("phases", Phase * MAXPHASES),
    ]

    def copy(self) -> "Config":
        c = Config()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(Config))
        return c


class DivInfo(C.Structure):
    _fields_ = [("np", C.c_uint32), ("npb", C.c_uint32), ("npbok", C.c_uint32),
                ("npbper", C.c_uint32), ("npfper", C.c_uint32),
                ("npbout", C.c_uint32), ("npfout", C.c_uint32), ("npbignore", C.c_uint32),
                ("npboutignore", C.c_uint32), ("npfoutignore", C.c_uint32),
                ("ncx", C.c_uint32), ("ncy", C.c_uint32), ("ncz", C.c_uint32), ("nct", C.c_uint32),
                ("cellmin", C.c_uint32 * 3), ("cellmax", C.c_uint32 * 3),
                ("domcells", C.c_uint32 * 3), ("domcellcode", C.c_uint32),
                ("dividefull", C.c_uint32), ("nctt", C.c_uint32),
                ("scell", C.c_double), ("domposmin", C.c_double * 3)]


class ForceRes(C.Structure):
    _fields_ = [("viscdtmax", C.c_float), ("acemax", C.c_float), ("velmax", C.c_float), ("viscetadtmax", C.c_float),
                ("dt1", C.c_double), ("dt2", C.c_double), ("dt3", C.c_double), ("dt", C.c_double)]


class MotionData(C.Structure):
    """dsph_motiondata = StMotionData (S/JDsMotionDef.h:33-44)."""
    _fields_ = [("idbegin", C.c_uint32), ("count", C.c_uint32), ("type", C.c_int32), ("reserved", C.c_int32),
                ("linmov", C.c_double * 3), ("linvel", C.c_double * 3), ("matmov", C.c_double * 16)]


class Timers(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("nl_limits", "nl_presort", "nl_radixsort", "nl_cellbegin", "nl_sortdata",
                                           "nl_outcheck", "cf_preforces", "cf_forces", "su_computestep", "su_periodic",
                                           "su_shifting", "total")] + [("kernel_launches", C.c_uint64)]


class Case(C.Structure):
    _fields_ = [("cfg", Config), ("np", C.c_uint32), ("npb", C.c_uint32),
                ("idp", C.POINTER(C.c_uint32)), ("pos_xyz", C.POINTER(C.c_double)), ("velrhop_xyzw", C.POINTER(C.c_float)),
                ("caseposmin", C.c_double * 3), ("caseposmax", C.c_double * 3), ("name", C.c_char * 64),
                ("nfluidblocks", C.c_uint32), ("fluidblock_count", C.c_uint32 * MAXPHASES)]


class Gauge(C.Structure):
    """dsph_gauge (include/dsph_b200.h): one in-solver gauge."""
    _fields_ = [("type", C.c_int32), ("output", C.c_int32), ("name", C.c_char * 32),
                ("computestart", C.c_double), ("computeend", C.c_double), ("computedt", C.c_double),
                ("outputstart", C.c_double), ("outputend", C.c_double), ("outputdt", C.c_double),
                ("point0", C.c_double * 3), ("point2", C.c_double * 3), ("pointdp", C.c_double), ("masslimit", C.c_float),
                ("distlimit", C.c_float), ("height", C.c_double),
                ("mkbound", C.c_uint32), ("idbegin", C.c_uint32), ("count", C.c_uint32), ("code", C.c_uint16), ("pad_", C.c_uint16)]


class GaugeRes(C.Structure):
    _fields_ = [("timestep", C.c_double), ("value", C.c_float * 3), ("computed", C.c_int32)]


GAUGE_VEL, GAUGE_SWL, GAUGE_MAXZ, GAUGE_FORCE = 1, 2, 3, 4


class SphError(RuntimeError):
    pass


def _ptr(a, ctype):
    return None if a is None else a.ctypes.data_as(C.POINTER(ctype))


class SphLib:
    """One shared library exporting <prefix>create/particles_upload/celldiv_run/forces_run/step_* ..."""

    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise SphError(f"shared library not found: {path} (run __graft_entry__.build())")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        f = self._f
        vp = C.c_void_p
        f("last_error", C.c_char_p)
        if prefix == "dsph_":
            f("ctx_create", C.c_int, C.POINTER(vp), C.c_int, C.POINTER(Config), C.c_uint32)
            f("ctx_destroy", C.c_int, vp)
        else:
            f("create", C.c_int, C.POINTER(vp), C.POINTER(Config))
            f("destroy", C.c_int, vp)
        f("particles_upload", C.c_int, vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16),
          C.POINTER(C.c_double), C.POINTER(C.c_float))
        f("particles_set_velrhopm1", C.c_int, vp, C.POINTER(C.c_float))
        f("particles_set_spstau", C.c_int, vp, C.POINTER(C.c_float))
        f("particles_download", C.c_int, vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint16),
          C.POINTER(C.c_uint32), C.POINTER(C.c_double), C.POINTER(C.c_float))
        f("np", C.c_uint32, vp)
        f("npb", C.c_uint32, vp)
        f("celldiv_run", C.c_int, vp, C.c_int, C.POINTER(DivInfo))
        f("celldiv_get_begincell", C.c_int, vp, C.POINTER(C.c_uint32), C.c_uint32)
        f("forces_run", C.c_int, vp, C.c_int, C.POINTER(ForceRes))
        f("forces_download", C.c_int, vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float))
        f("step_verlet", C.c_int, vp, C.c_double)
        f("step_sympre", C.c_int, vp, C.c_double)
        f("step_symcor", C.c_int, vp, C.c_double)
        f("step", C.c_int, vp, C.POINTER(C.c_double), C.POINTER(ForceRes))
        f("symplectic_dtpre", C.c_double, vp)
        f("set_symplectic_dtpre", C.c_int, vp, C.c_double)
        f("pips", C.c_int, vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64))
        # x-slab multi-GPU
        f("mgpu_config", C.c_int, vp, C.c_int, C.c_int, C.c_uint32, C.c_uint32)
        f("mgpu_record_bytes", C.c_uint32, vp)
        f("mgpu_count", C.c_int, vp, C.POINTER(C.c_uint32))
        f("mgpu_pack", C.c_int, vp, C.c_void_p, C.c_void_p)
        f("mgpu_unpack", C.c_int, vp, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p)
        if prefix == "dsph_":                                   # native transport (NCCL inside the library, mgpu_comm.cu)
            f("mgpu_unique_id", C.c_int, C.c_void_p)
            f("mgpu_nccl_version", C.c_int)
            f("mgpu_comm_init", C.c_int, vp, C.c_void_p, C.c_int, C.c_int)
            f("mgpu_comm_attach", C.c_int, vp, C.c_void_p, C.c_int, C.c_int)
            f("mgpu_config_cuts", C.c_int, vp, C.c_int, C.c_int, C.POINTER(C.c_uint32))
            f("mgpu_halo_cells", C.c_uint32, vp)
            f("mgpu_exchange", C.c_int, vp)
            f("mgpu_celldiv", C.c_int, vp, C.POINTER(DivInfo))
            f("mgpu_forces", C.c_int, vp, C.c_int, C.POINTER(ForceRes))
            f("mgpu_step", C.c_int, vp, C.POINTER(C.c_double), C.POINTER(ForceRes))
            f("mgpu_rebalance", C.c_int, vp, C.c_double, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_uint32))
            f("mgpu_measured_cost", C.c_double, vp)
            f("mgpu_profile", C.c_int, vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64))
            f("mgpu_sent_records", C.c_int, vp, C.POINTER(C.c_uint64))
            f("mgpu_transport", C.c_int, vp, C.POINTER(C.c_int), C.POINTER(C.c_uint64))
            f("mgpu_balanced_cuts", C.c_int, C.POINTER(C.c_double), C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_uint32))
            f("celldiv_last_info", C.c_int, vp, C.POINTER(DivInfo))
        if prefix == "dsph_":
            f("set_mdbc_threshold", C.c_int, vp, C.c_float)
            f("case_normals", C.c_int, C.c_char_p, C.c_char_p, C.POINTER(C.c_double), C.c_uint32, C.POINTER(C.c_uint32))
            f("case_gauges", C.c_int, C.c_char_p, C.c_char_p, C.POINTER(Gauge), C.c_uint32, C.POINTER(C.c_uint32))
        f("motion_run", C.c_int, vp, C.POINTER(MotionData), C.c_uint32, C.c_double)
        f("gauges_config", C.c_int, vp, C.c_uint32, C.POINTER(Gauge))
        f("gauges_run", C.c_int, vp, C.c_double, C.c_int, C.POINTER(GaugeRes))
        f("gauges_results", C.c_int, vp, C.POINTER(GaugeRes))
        f("gauges_set_time", C.c_int, vp, C.c_double)
        f("particles_set_boundnormals", C.c_int, vp, C.POINTER(C.c_float), C.c_uint32)
        f("particles_get_boundnormals", C.c_int, vp, C.POINTER(C.c_float), C.c_uint32)
        if prefix == "dsph_":
            f("case_motion_open", C.c_int, C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(C.c_uint32))
            f("case_motion_process", C.c_int, vp, C.c_double, C.c_double, C.POINTER(MotionData), C.POINTER(C.c_int))
            f("case_motion_close", C.c_int, vp)
            f("motion_attach", C.c_int, vp, vp)
        f("forces_launch", C.c_int, vp, C.c_int)
        f("forces_maxima_ptr", C.c_void_p, vp)
        f("forces_finish", C.c_int, vp, C.POINTER(ForceRes))
        f("domain_info", C.c_int, C.POINTER(Config), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32))

    def _f(self, name, restype, *argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype, fn.argtypes = restype, list(argtypes)
        return fn

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def check(self, rc):
        if rc != 0:
            msg = self.fn("last_error")()
            raise SphError(f"{self.prefix}* call failed ({rc}): {msg.decode(errors='replace') if msg else ''}")


class SphContext:
    """Particle state + hot path of one solver instance (the JSphGpuSingle work-alike seen through the C ABI)."""

    def __init__(self, lib: SphLib, cfg: Config, np_capacity: int = 0, device: int = 0):
        self.lib, self.cfg = lib, cfg
        self.h = C.c_void_p()
        if lib.prefix == "dsph_":
            lib.check(lib.fn("ctx_create")(C.byref(self.h), device, C.byref(cfg), np_capacity))
        else:
            lib.check(lib.fn("create")(C.byref(self.h), C.byref(cfg)))

    def close(self):
        if self.h:
            self.lib.fn("ctx_destroy" if self.lib.prefix == "dsph_" else "destroy")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- state
    def upload(self, idp, pos, velrhop, npb, code=None):
        idp = np.ascontiguousarray(idp, np.uint32)
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        velrhop = np.ascontiguousarray(velrhop, np.float32).reshape(-1, 4)
        code = None if code is None else np.ascontiguousarray(code, np.uint16)
        n = idp.shape[0]
        assert pos.shape[0] == n and velrhop.shape[0] == n
        self.lib.check(self.lib.fn("particles_upload")(self.h, n, int(npb), _ptr(idp, C.c_uint32), _ptr(code, C.c_uint16),
                                                       _ptr(pos, C.c_double), _ptr(velrhop, C.c_float)))

    def set_velrhopm1(self, v):
        v = np.ascontiguousarray(v, np.float32).reshape(-1, 4)
        self.lib.check(self.lib.fn("particles_set_velrhopm1")(self.h, _ptr(v, C.c_float)))

    def set_boundnormals(self, n):
        """mDBC: vectors particle -> ghost node (xyz) of the first len(n) particles of the uploaded order (boundary particles first)."""
        n = np.ascontiguousarray(n, np.float32).reshape(-1, 3)
        self.lib.check(self.lib.fn("particles_set_boundnormals")(self.h, _ptr(n, C.c_float), n.shape[0]))

    def boundnormals(self, count):
        n = np.zeros((int(count), 3), np.float32)
        self.lib.check(self.lib.fn("particles_get_boundnormals")(self.h, _ptr(n, C.c_float), int(count)))
        return n

    def set_spstau(self, t):
        t = np.ascontiguousarray(t, np.float32).reshape(-1, 6)
        self.lib.check(self.lib.fn("particles_set_spstau")(self.h, _ptr(t, C.c_float)))

    @property
    def np_(self):
        return int(self.lib.fn("np")(self.h))

    @property
    def npb(self):
        return int(self.lib.fn("npb")(self.h))

    def download(self, first=0, count=None):
        n = self.np_ - first if count is None else count
        idp = np.empty(n, np.uint32); code = np.empty(n, np.uint16); dcell = np.empty(n, np.uint32)
        pos = np.empty((n, 3), np.float64); vr = np.empty((n, 4), np.float32)
        self.lib.check(self.lib.fn("particles_download")(self.h, first, n, _ptr(idp, C.c_uint32), _ptr(code, C.c_uint16),
                                                         _ptr(dcell, C.c_uint32), _ptr(pos, C.c_double), _ptr(vr, C.c_float)))
        return dict(idp=idp, code=code, dcell=dcell, pos=pos, velrhop=vr)

    # -- hot path
    def celldiv(self, updateperiodic=True) -> DivInfo:
        info = DivInfo()
        self.lib.check(self.lib.fn("celldiv_run")(self.h, 1 if updateperiodic else 0, C.byref(info)))
        self.divinfo = info
        return info

    def begincell(self):
        n = int(self.divinfo.nctt)
        b = np.empty(n, np.uint32)
        self.lib.check(self.lib.fn("celldiv_get_begincell")(self.h, _ptr(b, C.c_uint32), n))
        return b

    def forces(self, interstep=INTERSTEP_SYMPREDICTOR) -> ForceRes:
        r = ForceRes()
        self.lib.check(self.lib.fn("forces_run")(self.h, interstep, C.byref(r)))
        return r

    def forces_download(self, first=0, count=None):
        n = self.np_ - first if count is None else count
        ace = np.empty((n, 3), np.float32); ar = np.empty(n, np.float32)
        self.lib.check(self.lib.fn("forces_download")(self.h, first, n, _ptr(ace, C.c_float), _ptr(ar, C.c_float)))
        return ace, ar

    def step_verlet(self, dt):
        self.lib.check(self.lib.fn("step_verlet")(self.h, float(dt)))

    def step_sympre(self, dt):
        self.lib.check(self.lib.fn("step_sympre")(self.h, float(dt)))

    def step_symcor(self, dt):
        self.lib.check(self.lib.fn("step_symcor")(self.h, float(dt)))

    def step(self):
        dt = C.c_double(); r = ForceRes()
        self.lib.check(self.lib.fn("step")(self.h, C.byref(dt), C.byref(r)))
        return dt.value, r

    @property
    def symplectic_dtpre(self):
        return float(self.lib.fn("symplectic_dtpre")(self.h))

    @symplectic_dtpre.setter
    def symplectic_dtpre(self, v):
        self.lib.check(self.lib.fn("set_symplectic_dtpre")(self.h, float(v)))

    # -- moving boundaries
    def motion_run(self, objs, stepdt):
        """RunMotion with the motion data of this step (list of MotionData)."""
        arr = (MotionData * max(1, len(objs)))(*objs)
        self.lib.check(self.lib.fn("motion_run")(self.h, arr, len(objs), float(stepdt)))

    def motion_attach(self, motion):
        """Let step() evaluate and apply the motion laws of a CaseMotion (None detaches)."""
        self._motion = motion
        self.lib.check(self.lib.fn("motion_attach")(self.h, motion.h if motion is not None else None))

    # -- in-solver gauges
    def gauges_config(self, gauges):
        arr = (Gauge * len(gauges))(*gauges)
        self._ngauges = len(gauges)
        self.lib.check(self.lib.fn("gauges_config")(self.h, len(gauges), arr))

    def gauges_run(self, timestep, force_all=False):
        res = (GaugeRes * self._ngauges)()
        self.lib.check(self.lib.fn("gauges_run")(self.h, float(timestep), 1 if force_all else 0, res))
        return list(res)

    def gauges_results(self):
        res = (GaugeRes * self._ngauges)()
        self.lib.check(self.lib.fn("gauges_results")(self.h, res))
        return list(res)

    def gauges_set_time(self, t):
        self.lib.check(self.lib.fn("gauges_set_time")(self.h, float(t)))

    # -- x-slab multi-GPU (see mgpu.py)
    def measure_points(self, points):
        """Pressure, velocity and kernel sum at `points` [k,3] (Shepard-corrected interpolation over the fluid particles)."""
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        n = pts.shape[0]
        press = np.zeros(n, np.float32); vel = np.zeros((n, 3), np.float32); kc = np.zeros(n, np.float32)
        fn = self.lib.fn("measure_points")
        fn.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        self.lib.check(fn(self.h, n, pts.ctypes.data_as(C.POINTER(C.c_double)), press.ctypes.data_as(C.POINTER(C.c_float)),
                          vel.ctypes.data_as(C.POINTER(C.c_float)), kc.ctypes.data_as(C.POINTER(C.c_float))))
        return press, vel, kc

    def mgpu_config(self, rank, nranks, xlo, xhi):
        self.lib.check(self.lib.fn("mgpu_config")(self.h, int(rank), int(nranks), int(xlo), int(xhi)))

    def mgpu_count(self):
        """[left: bound ghosts, bound migrants, fluid migrants, fluid ghosts, right: the same four]"""
        c = (C.c_uint32 * 8)()
        self.lib.check(self.lib.fn("mgpu_count")(self.h, c))
        return [int(v) for v in c]

    def mgpu_column_counts(self, ncol):
        """Owned particles per cell column (fluid, boundary): input of the slab re-balancing."""
        f = np.zeros(ncol, np.uint32); b = np.zeros(ncol, np.uint32)
        fn = self.lib.fn("mgpu_column_counts")
        fn.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_uint32]
        self.lib.check(fn(self.h, f.ctypes.data_as(C.POINTER(C.c_uint32)), b.ctypes.data_as(C.POINTER(C.c_uint32)), int(ncol)))
        return f, b

    def mgpu_pack(self, ptr_left, ptr_right):
        self.lib.check(self.lib.fn("mgpu_pack")(self.h, ptr_left, ptr_right))

    def mgpu_unpack(self, counts, ptr_left, ptr_right):
        c = (C.c_uint32 * 8)(*[int(v) for v in counts])
        self.lib.check(self.lib.fn("mgpu_unpack")(self.h, c, ptr_left, ptr_right))

    def forces_launch(self, interstep=INTERSTEP_SYMPREDICTOR):
        self.lib.check(self.lib.fn("forces_launch")(self.h, interstep))

    def forces_maxima_ptr(self):
        return int(self.lib.fn("forces_maxima_ptr")(self.h))

    def forces_finish(self) -> ForceRes:
        r = ForceRes()
        self.lib.check(self.lib.fn("forces_finish")(self.h, C.byref(r)))
        return r

    def domain_info(self):
        sc = C.c_double(); pm = (C.c_double * 3)(); dc = (C.c_uint32 * 3)()
        self.lib.check(self.lib.fn("domain_info")(C.byref(self.cfg), C.byref(sc), pm, dc))
        return sc.value, [float(v) for v in pm], [int(v) for v in dc]

    def pips(self):
        v = [C.c_uint64() for _ in range(4)]
        self.lib.check(self.lib.fn("pips")(self.h, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)
