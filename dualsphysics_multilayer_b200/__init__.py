"""dualsphysics_multilayer_b200 -- B200-native per-step particle hot path of DualSPHysics (cell-linked list ->
Interaction_Forces -> ComputeStep) behind the C ABI of ``include/dsph_b200.h``.

The compute path is hand-written CUDA for sm_100a inside ``libdsph_b200.so`` (built in-tree by
``__graft_entry__.build()``).  This Python layer is only the test/bench harness view of that C ABI (ctypes); the
reference is compiled C++, so the host driver that mirrors JSph/JSphGpuSingle lives in ``csrc/host`` (C++).
There is no CPU fallback: if the library is missing or no CUDA device exists, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .capi import (Case, Config, DivInfo, ForceRes, SphContext, SphError, SphLib, Timers)  # noqa: F401
from . import capi  # noqa: F401

_PKGDIR = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(_PKGDIR, "libdsph_b200.so")
_lib = None


def lib() -> SphLib:
    """The product shared library (loaded once).  Raises SphError when it has not been built."""
    global _lib
    if _lib is None:
        L = SphLib(LIBPATH, "dsph_")
        l = L.lib
        l.dsph_version.restype = C.c_char_p
        l.dsph_device_count.restype = C.c_int
        l.dsph_case_make.restype = C.c_int
        l.dsph_case_make.argtypes = [C.POINTER(Case), C.c_char_p, C.c_uint64, C.c_uint32]
        l.dsph_case_make_slab.restype = C.c_int
        l.dsph_case_make_slab.argtypes = [C.POINTER(Case), C.c_char_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        l.dsph_case_free.argtypes = [C.POINTER(Case)]
        l.dsph_case_load.restype = C.c_int
        l.dsph_case_load.argtypes = [C.POINTER(Case), C.c_char_p, C.c_char_p]
        l.dsph_case_save.restype = C.c_int
        l.dsph_case_save.argtypes = [C.POINTER(Case), C.c_char_p, C.c_char_p]
        l.dsph_run_case.restype = C.c_int
        l.dsph_run_case.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]
        l.dsph_get_timers.restype = C.c_int
        l.dsph_get_timers.argtypes = [C.c_void_p, C.POINTER(Timers), C.c_int]
        l.dsph_set_force_variant.argtypes = [C.c_void_p, C.c_int]
        l.dsph_set_verletstep.argtypes = [C.c_void_p, C.c_int]
        l.dsph_get_force_variant.restype = C.c_int
        l.dsph_get_force_variant.argtypes = [C.c_void_p]
        l.dsph_last_force_ms.restype = C.c_double
        l.dsph_last_force_ms.argtypes = [C.c_void_p]
        l.dsph_cuda_stream.restype = C.c_void_p
        l.dsph_cuda_stream.argtypes = [C.c_void_p]
        l.dsph_constants_upload.restype = C.c_int
        l.dsph_constants_upload.argtypes = [C.c_void_p, C.POINTER(Config)]
        _lib = L
    return _lib


def device_count() -> int:
    return int(lib().lib.dsph_device_count())


class CaseData:
    """Host copy of a case: config + particle arrays (numpy)."""

    def __init__(self, cfg: Config, idp, pos, velrhop, npb, name="Case", caseposmin=None, caseposmax=None,
                 fluidblocks=None):
        self.cfg, self.idp, self.pos, self.velrhop, self.npb, self.name = cfg, idp, pos, velrhop, int(npb), name
        self.caseposmin, self.caseposmax = caseposmin, caseposmax
        self.fluidblocks = fluidblocks or [int(cfg.casenfluid)]

    @property
    def np_(self):
        return int(self.idp.shape[0])

    def codes(self):
        """typecode per particle: fixed block 0, fluid blocks by rank (JSphMk::GetCodeById)."""
        code = np.full(self.np_, capi.CODE_TYPE_FIXED, np.uint16)
        nb = int(self.cfg.casenfixed)
        nmov = int(self.cfg.casenmoving)
        if nmov:                                         # one moving block (the synthetic cases have one plate)
            code[(self.idp >= nb) & (self.idp < nb + nmov)] = capi.CODE_TYPE_MOVING
        start = nb + nmov
        for rank, cnt in enumerate(self.fluidblocks):
            sel = (self.idp >= start) & (self.idp < start + cnt)
            code[sel] = capi.CODE_TYPE_FLUID | rank
            start += cnt
        return code


class CaseMotion:
    """Motion laws of <case><execution><motion> evaluated by the library's host code (dsph_case_motion_*), JDsMotion work-alike."""

    def __init__(self, directory, casename="Case"):
        L = lib()
        self._lib = L
        self.h = C.c_void_p()
        n = C.c_uint32()
        L.check(L.fn("case_motion_open")(str(directory).encode(), casename.encode(), C.byref(self.h), C.byref(n)))
        self.nobj = int(n.value)

    def process(self, timestep, dt):
        """= JDsMotion::ProcesTime(simple, timestep, dt) + GetMotionData of every object -> (list of MotionData, active)."""
        out = (capi.MotionData * max(1, self.nobj))()
        act = C.c_int()
        self._lib.check(self._lib.fn("case_motion_process")(self.h, float(timestep), float(dt), out, C.byref(act)))
        return list(out)[: self.nobj], bool(act.value)

    def close(self):
        if self.h:
            self._lib.fn("case_motion_close")(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _case_from_c(c: Case) -> CaseData:
    n = int(c.np)
    idp = np.ctypeslib.as_array(c.idp, shape=(n,)).copy()
    pos = np.ctypeslib.as_array(c.pos_xyz, shape=(n, 3)).copy()
    vr = np.ctypeslib.as_array(c.velrhop_xyzw, shape=(n, 4)).copy()
    nfb = int(c.nfluidblocks)
    blocks = [int(c.fluidblock_count[i]) for i in range(nfb)] if nfb else None
    cd = CaseData(c.cfg.copy(), idp, pos, vr, int(c.npb), c.name.decode(), list(c.caseposmin), list(c.caseposmax), blocks)
    lib().lib.dsph_case_free(C.byref(c))
    return cd


def make_case(kind: str, np_target: int, seed: int = 12345) -> CaseData:
    """Synthetic benchmark cases of SURVEY.md 8(d): 'dambreak', 'periodicbox', 'nn2layer'."""
    c = Case()
    L = lib()
    L.check(L.lib.dsph_case_make(C.byref(c), kind.encode(), int(np_target), int(seed)))
    return _case_from_c(c)


def make_case_slab(kind: str, np_target: int, cx_lo: int, cx_hi: int, seed: int = 12345) -> CaseData:
    """The particles of cell columns [cx_lo,cx_hi) of make_case(kind, np_target); cfg/ids are those of the full case."""
    c = Case()
    L = lib()
    L.check(L.lib.dsph_case_make_slab(C.byref(c), kind.encode(), int(np_target), int(seed), int(cx_lo), int(cx_hi)))
    return _case_from_c(c)


def load_case(dirname: str, casename: str) -> CaseData:
    """Reads <dirname>/<casename>.xml + .bi4 (GenCase output format)."""
    c = Case()
    L = lib()
    L.check(L.lib.dsph_case_load(C.byref(c), dirname.encode(), casename.encode()))
    return _case_from_c(c)


def case_normals(dirname: str, casename: str):
    """mDBC normal vectors (particle -> boundary limit) of <casename>_Normals.nbi4, one row per boundary particle in id order
    (empty array when the case has no such file)."""
    L = lib()
    n = C.c_uint32()
    L.check(L.fn("case_normals")(dirname.encode(), casename.encode(), None, 0, C.byref(n)))
    out = np.zeros((n.value, 3), np.float64)
    if n.value:
        L.check(L.fn("case_normals")(dirname.encode(), casename.encode(), out.ctypes.data_as(C.POINTER(C.c_double)), n.value, C.byref(n)))
    return out


def case_gauges(dirname: str, casename: str):
    """Gauge definitions of <special><gauges> in <dirname>/<casename>.xml (list of capi.Gauge)."""
    L = lib()
    n = C.c_uint32()
    out = (capi.Gauge * 256)()
    L.check(L.fn("case_gauges")(dirname.encode(), casename.encode(), out, 256, C.byref(n)))
    return [out[i] for i in range(min(n.value, 256))]


def load_part(filename: str):
    """Reads one PART file (Part_NNNN.bi4 of this driver or of the reference).  Returns (CaseData, timestep, step)."""
    c = Case()
    L = lib()
    t = C.c_double(); st = C.c_uint32()
    L.lib.dsph_part_load.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint32)]
    L.check(L.lib.dsph_part_load(C.byref(c), filename.encode(), C.byref(t), C.byref(st)))
    return _case_from_c(c), t.value, st.value


def _case_to_c(cd: CaseData):
    """ctypes view of a CaseData (the numpy arrays are returned too: they must outlive the call)."""
    c = Case()
    c.cfg = cd.cfg
    c.np, c.npb = cd.np_, cd.npb
    idp = np.ascontiguousarray(cd.idp, np.uint32)
    pos = np.ascontiguousarray(cd.pos, np.float64)
    vr = np.ascontiguousarray(cd.velrhop, np.float32)
    c.idp = idp.ctypes.data_as(C.POINTER(C.c_uint32))
    c.pos_xyz = pos.ctypes.data_as(C.POINTER(C.c_double))
    c.velrhop_xyzw = vr.ctypes.data_as(C.POINTER(C.c_float))
    for i in range(3):
        c.caseposmin[i] = cd.caseposmin[i]
        c.caseposmax[i] = cd.caseposmax[i]
    c.name = cd.name.encode()
    c.nfluidblocks = len(cd.fluidblocks)
    for i, n in enumerate(cd.fluidblocks):
        c.fluidblock_count[i] = n
    return c, (idp, pos, vr)


def save_partset(cd: CaseData, dirout: str, cpart: int, timestep: float, step: int = 0, symplecticdtpre: float = 0.0, piecemax: int = 0,
                 out=None) -> None:
    """Host only: the PART output set of one instant (Part_Head.ibi4, Part_NNNN.bi4 or its pieces, PartInfo.ibi4, PartOut_000.obi4).
    out = (idp, pos[n,3], velrhop[n,4], motive) of the excluded particles, or None."""
    c, keep = _case_to_c(cd)
    os.makedirs(dirout, exist_ok=True)
    L = lib()
    fn = L.lib.dsph_partset_save
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(Case), C.c_char_p, C.c_uint32, C.c_double, C.c_uint32, C.c_double, C.c_uint32, C.c_uint32,
                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    if out is None:
        L.check(fn(C.byref(c), dirout.encode(), cpart, timestep, step, symplecticdtpre, piecemax, 0, None, None, None, None))
    else:
        oid = np.ascontiguousarray(out[0], np.uint32); opos = np.ascontiguousarray(out[1], np.float64)
        ovr = np.ascontiguousarray(out[2], np.float32); omo = np.ascontiguousarray(out[3], np.uint8)
        L.check(fn(C.byref(c), dirout.encode(), cpart, timestep, step, symplecticdtpre, piecemax, len(oid),
                   oid.ctypes.data, opos.ctypes.data, ovr.ctypes.data, omo.ctypes.data))


def load_partset(dirname: str, cpart: int):
    """Every piece of PART `cpart` in `dirname`, concatenated.  Returns (CaseData, timestep, step, npiece)."""
    c = Case()
    L = lib()
    t = C.c_double(); st = C.c_uint32(); npc = C.c_uint32()
    fn = L.lib.dsph_partset_load
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.check(fn(C.byref(c), dirname.encode(), cpart, C.byref(t), C.byref(st), C.byref(npc)))
    return _case_from_c(c), t.value, st.value, npc.value


def save_case(cd: CaseData, dirname: str, casename: str) -> None:
    c = Case()
    c.cfg = cd.cfg
    c.np, c.npb = cd.np_, cd.npb
    idp = np.ascontiguousarray(cd.idp, np.uint32)
    pos = np.ascontiguousarray(cd.pos, np.float64)
    vr = np.ascontiguousarray(cd.velrhop, np.float32)
    c.idp = idp.ctypes.data_as(C.POINTER(C.c_uint32))
    c.pos_xyz = pos.ctypes.data_as(C.POINTER(C.c_double))
    c.velrhop_xyzw = vr.ctypes.data_as(C.POINTER(C.c_float))
    for i in range(3):
        c.caseposmin[i] = cd.caseposmin[i]
        c.caseposmax[i] = cd.caseposmax[i]
    c.nfluidblocks = len(cd.fluidblocks)
    for i, n in enumerate(cd.fluidblocks):
        c.fluidblock_count[i] = n
    os.makedirs(dirname, exist_ok=True)
    L = lib()
    L.check(L.lib.dsph_case_save(C.byref(c), dirname.encode(), casename.encode()))


def create_context(cd_or_cfg, np_capacity: int = 0, device: int = 0) -> SphContext:
    """GPU context (JSphGpuSingle work-alike).  Raises SphError without a CUDA device."""
    cfg = cd_or_cfg.cfg if isinstance(cd_or_cfg, CaseData) else cd_or_cfg
    return SphContext(lib(), cfg, np_capacity, device)


class _RunOpts(C.Structure):
    _fields_ = [("dir", C.c_char_p), ("casename", C.c_char_p), ("dirout", C.c_char_p),
                ("device", C.c_int32), ("nsteps", C.c_int32), ("savepart", C.c_int32), ("forcevariant", C.c_int32),
                ("timemax", C.c_double), ("timeout", C.c_double), ("restart", C.c_int32), ("partbegin", C.c_int32),
                ("partbegindir", C.c_char_p), ("mgpu", C.c_int32), ("rebalancesteps", C.c_int32), ("boundweight", C.c_double),
                ("piecemax", C.c_uint32), ("reserved_", C.c_uint32)]


def run_case(dirname: str, casename: str, dirout: str, device: int = 0, nsteps: int = 0, savepart: bool = True,
             timemax: float = -1.0, timeout: float = -1.0, partbegin: int = -1, partbegindir: str = None,
             mgpu: int = 0, rebalancesteps: int = 0, piecemax: int = 0) -> None:
    """The stand-alone driver (dsph_run_case_ex): Case.xml + Case.bi4 in, Part_Head.ibi4 / Part_NNNN.bi4 (or pieces) /
    PartInfo.ibi4 / PartOut_000.obi4 / Run.out / Run.csv out.  mgpu > 1: x-slab run on that many GPUs of this process."""
    L = lib()
    o = _RunOpts(dirname.encode(), casename.encode(), dirout.encode(), device, nsteps, 1 if savepart else 0, 0,
                 timemax, timeout, 1 if partbegin >= 0 else 0, max(partbegin, 0), partbegindir.encode() if partbegindir else None,
                 mgpu, rebalancesteps, 0.0, piecemax, 0)
    L.lib.dsph_run_case_ex.argtypes = [C.POINTER(_RunOpts)]
    L.check(L.lib.dsph_run_case_ex(C.byref(o)))
