"""x-slab domain decomposition over the GPUs of one box: one process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on GPUs; gloo on CPU for the tests, where the per-rank engine is the CPU checker).

Per neighbour-list build every rank
  1. classifies its particles (`mgpu_count`): fluid particles whose cell column left the slab migrate to the neighbour,
     particles within 2h (ScellDiv cells) of a cut are copied to the neighbour as ghosts (CODE_PERIODIC: neighbours only),
  2. exchanges the six counts and then the packed 88-byte records with its <=2 neighbours (send/recv),
  3. appends what it received (`mgpu_unpack`) and runs the ordinary cell division.
Per force evaluation the three maxima (viscdt, |ace|^2, |vel|^2: uint32 bit patterns of non-negative floats, so integer
MAX is float MAX) are all-reduced before the time step is derived, so every rank uses the same dt.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch
import torch.distributed as dist

from . import capi


def balanced_cuts(colcount: np.ndarray, nranks: int, minwidth: int) -> list:
    """Cut positions (cell columns) that minimise the work of the busiest slab; slabs keep at least `minwidth` columns.
    Returns nranks+1 increasing column indices, first 0 and last len(colcount) (dsph_mgpu_balanced_cuts: host code of the library)."""
    col = np.ascontiguousarray(colcount, dtype=np.float64)
    if nranks * minwidth > len(col):
        raise ValueError("too many ranks for %d cell columns (minimum slab width %d)" % (len(col), minwidth))
    from . import lib as _getlib
    lib = _getlib()
    out = (C.c_uint32 * (nranks + 1))()
    lib.check(lib.fn("mgpu_balanced_cuts")(col.ctypes.data_as(C.POINTER(C.c_double)), len(col), int(nranks), int(minwidth), out))
    return [int(v) for v in out]


class _CudaView:
    """Minimal __cuda_array_interface__ wrapper so torch can view a raw device pointer owned by the C library."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class SlabDriver:
    def __init__(self, ctx: capi.SphContext, rank: int, world: int, cuts, device: str, group=None):
        self.ctx, self.rank, self.world, self.group = ctx, rank, world, group
        self.device = torch.device(device)
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None
        self.cuts = list(cuts)
        ctx.mgpu_config(rank, world, self.cuts[rank], self.cuts[rank + 1])
        self.recbytes = int(ctx.lib.fn("mgpu_record_bytes")(ctx.h))
        self._buf = {}
        ptr = ctx.forces_maxima_ptr()
        if self.device.type == "cuda":
            self.maxima = torch.as_tensor(_CudaView(ptr, 4, "<i4"), device=self.device)
            self._libstream = torch.cuda.ExternalStream(int(ctx.lib.lib.dsph_cuda_stream(ctx.h)), device=self.device)
        else:
            self.maxima = torch.from_numpy(np.ctypeslib.as_array((C.c_int32 * 4).from_address(ptr)))
        self.sent_records = 0
        self.prof = None                      # set to {} to accumulate wall seconds per phase (adds device syncs)
        self.symplectic = (ctx.cfg.tstep == capi.STEP_SYMPLECTIC)

    # ---- buffers -------------------------------------------------------------------------------------------------
    def _buffer(self, name, nrec):
        need = max(1, nrec) * self.recbytes
        b = self._buf.get(name)
        if b is None or b.numel() < need:
            b = torch.empty(int(need * 1.25) + 4096, dtype=torch.uint8, device=self.device)
            self._buf[name] = b
        return b

    # ---- halo + migration exchange ---------------------------------------------------------------------------------
    def exchange(self):
        if self.world == 1:
            return
        ctx = self.ctx
        t0 = self._tick()
        counts = ctx.mgpu_count()                                   # [L: bg, bm, fm, fg, R: bg, bm, fm, fg]
        t0 = self._tick("xch_count", t0)
        mine = torch.tensor(counts, dtype=torch.int64, device=self.device)
        fromleft = torch.zeros(4, dtype=torch.int64, device=self.device)
        fromright = torch.zeros(4, dtype=torch.int64, device=self.device)
        toleft = mine[0:4].contiguous(); toright = mine[4:8].contiguous()
        ops = []
        if self.left is not None:
            ops += [dist.P2POp(dist.isend, toleft, self.left, self.group), dist.P2POp(dist.irecv, fromleft, self.left, self.group)]
        if self.right is not None:
            ops += [dist.P2POp(dist.isend, toright, self.right, self.group), dist.P2POp(dist.irecv, fromright, self.right, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        nl, nr = int(sum(counts[0:4])), int(sum(counts[4:8]))
        rl = [int(v) for v in fromleft.tolist()]; rr = [int(v) for v in fromright.tolist()]
        t0 = self._tick("xch_counts_comm", t0)
        sl, sr = self._buffer("sl", nl), self._buffer("sr", nr)
        ctx.mgpu_pack(sl.data_ptr(), sr.data_ptr())
        t0 = self._tick("xch_pack", t0)
        bl, br = self._buffer("rl", sum(rl)), self._buffer("rr", sum(rr))
        ops = []
        if self.left is not None:
            if nl:
                ops.append(dist.P2POp(dist.isend, sl[: nl * self.recbytes], self.left, self.group))
            if sum(rl):
                ops.append(dist.P2POp(dist.irecv, bl[: sum(rl) * self.recbytes], self.left, self.group))
        if self.right is not None:
            if nr:
                ops.append(dist.P2POp(dist.isend, sr[: nr * self.recbytes], self.right, self.group))
            if sum(rr):
                ops.append(dist.P2POp(dist.irecv, br[: sum(rr) * self.recbytes], self.right, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        t0 = self._tick("xch_records_comm", t0)
        ctx.mgpu_unpack(rl + rr, bl.data_ptr(), br.data_ptr())
        self._tick("xch_unpack", t0)
        self.sent_records += nl + nr

    def _tick(self, name=None, t0=0.0):
        if self.prof is None:
            return 0.0
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        t = time.perf_counter()
        if name:
            self.prof[name] = self.prof.get(name, 0.0) + (t - t0)
        return t

    # ---- re-balancing: move the cuts to the current particle distribution ------------------------------------------
    def rebalance(self, bound_weight=0.05, minwidth=None, rank_cost=None):
        """Collective.  New cuts from the global per-column histogram of owned particles (fluid 1, boundary bound_weight).
        A cut moves by less than the width of the narrower of its two slabs, so every particle changes owner by at most
        one rank; the next exchange migrates fluid AND boundary particles that are outside the new range and rebuilds the
        halos.  Returns True when the cuts changed.
        rank_cost (optional, one number per call and rank, e.g. `measured_cost()`): the measured cost of this rank's own
        work since the last call.  The columns of every slab are then scaled so that they add up to the slab's measured
        cost: the particle histogram only distributes the cost inside a slab, the measurement decides between slabs
        (free-surface and wall particles have fewer neighbours than interior ones, which no static weight captures)."""
        if self.world == 1:
            return False
        ctx = self.ctx
        ncol = int(ctx.domain_info()[2][0])
        fl, bd = ctx.mgpu_column_counts(ncol)
        w = torch.from_numpy(fl.astype(np.float64) + bound_weight * bd.astype(np.float64)).to(self.device)
        dist.all_reduce(w, op=dist.ReduceOp.SUM, group=self.group)
        hist = w.cpu().numpy()
        if rank_cost is not None:
            costs = torch.zeros(self.world, dtype=torch.float64, device=self.device)
            costs[self.rank] = float(rank_cost)
            dist.all_reduce(costs, op=dist.ReduceOp.SUM, group=self.group)
            costs = costs.cpu().numpy()
            if np.all(costs > 0):
                for r in range(self.world):
                    seg = hist[self.cuts[r]:self.cuts[r + 1]]
                    tot = seg.sum()
                    if tot > 0:
                        seg *= costs[r] / tot
        sd = 2 if int(ctx.cfg.cellmode) == 2 else 1          # DSPH_CELLMODE_HALF: cells of h, halo = 2 cells
        if minwidth is None:
            minwidth = 2 * sd + 1
        target = balanced_cuts(hist, self.world, minwidth)
        new = list(self.cuts)
        for r in range(1, self.world):
            room = min(self.cuts[r] - self.cuts[r - 1], self.cuts[r + 1] - self.cuts[r]) - 1
            new[r] = int(min(max(target[r], self.cuts[r] - room), self.cuts[r] + room))
        for r in range(1, self.world):                                 # keep every slab at least minwidth wide
            new[r] = max(new[r], new[r - 1] + minwidth)
        for r in range(self.world - 1, 0, -1):
            new[r] = min(new[r], new[r + 1] - minwidth)
        if new == self.cuts:
            return False
        self.cuts = new
        ctx.mgpu_config(self.rank, self.world, new[self.rank], new[self.rank + 1])
        return True

    def measured_cost(self):
        """Device milliseconds of this rank's own kernels (cell division, forces, updates: the CUDA-event timers of the
        context) since the last call; waiting for the neighbours is not part of it."""
        tm = capi.Timers()
        self.ctx.lib.lib.dsph_get_timers(self.ctx.h, C.byref(tm), 1)
        return (tm.nl_limits + tm.nl_presort + tm.nl_radixsort + tm.nl_cellbegin + tm.nl_sortdata + tm.nl_outcheck +
                tm.cf_preforces + tm.cf_forces + tm.su_computestep + tm.su_periodic + tm.su_shifting)

    # ---- the step pieces (same sequence as JSphGpuSingle::ComputeStep_Sym/_Ver + closing divide) -------------------
    def celldiv(self):
        self.exchange()
        t0 = self._tick()
        r = self.ctx.celldiv(False)
        self._tick("celldiv", t0)
        return r

    def forces(self, interstep):
        t0 = self._tick()
        self.ctx.forces_launch(interstep)
        t0 = self._tick("forces_kernels", t0)
        if self.world > 1:
            if self.device.type == "cuda":
                # the library launched on its own stream (dsph_cuda_stream): order the collective after the force kernels and
                # the read-back of dsph_forces_finish after the collective with stream waits, no device-wide synchronisation
                cur = torch.cuda.current_stream(self.device)
                cur.wait_stream(self._libstream)
                dist.all_reduce(self.maxima, op=dist.ReduceOp.MAX, group=self.group)
                self._libstream.wait_stream(cur)
            else:
                dist.all_reduce(self.maxima, op=dist.ReduceOp.MAX, group=self.group)
        t0 = self._tick("dt_allreduce", t0)
        return self.ctx.forces_finish()

    def step(self):
        ctx = self.ctx
        if self.symplectic:
            dt = ctx.symplectic_dtpre
            rp = self.forces(capi.INTERSTEP_SYMPREDICTOR)
            ctx.step_sympre(dt)
            self.celldiv()
            rc = self.forces(capi.INTERSTEP_SYMCORRECTOR)
            ctx.step_symcor(dt)
            ctx.symplectic_dtpre = min(rp.dt, rc.dt)
            res = rc
        else:
            res = self.forces(capi.INTERSTEP_VERLET)
            dt = res.dt
            ctx.step_verlet(dt)
        self.celldiv()
        return dt, res

    # ---- helpers -------------------------------------------------------------------------------------------------
    def owned(self):
        """Particles owned by this rank (ghosts removed), as downloaded arrays."""
        st = self.ctx.download()
        keep = (st["code"] & capi.CODE_MASKSPECIAL) == capi.CODE_NORMAL
        return {k: v[keep] for k, v in st.items()}


class NativeSlabDriver:
    """The same interface as SlabDriver over the library's native transport (csrc/mgpu_comm.cu): the exchange, the dt all-reduce,
    the step sequencing and the re-balancing run inside libdsph_b200.so with NCCL on the context's own stream -- this class only
    forwards.  torch.distributed is used once, to hand rank 0's 128-byte ncclUniqueId to the other ranks (any channel would do:
    bin/dsph_b200 -mgpu uses threads of one process)."""

    def __init__(self, ctx: capi.SphContext, rank: int, world: int, cuts, device: str = None, group=None, uid: bytes = None):
        self.ctx, self.rank, self.world = ctx, rank, world
        lib = ctx.lib
        if uid is None:
            box = [None]
            if rank == 0:
                buf = C.create_string_buffer(128)
                lib.check(lib.fn("mgpu_unique_id")(buf))
                box[0] = bytes(buf.raw)
            dist.broadcast_object_list(box, src=0, group=group)
            uid = box[0]
        self.cuts = [int(c) for c in cuts]
        lib.check(lib.fn("mgpu_comm_init")(ctx.h, C.c_char_p(uid), rank, world))
        self._set_cuts(self.cuts)
        self.recbytes = int(lib.fn("mgpu_record_bytes")(ctx.h))
        self.halo = int(lib.fn("mgpu_halo_cells")(ctx.h))
        self._prof_on = False

    def _set_cuts(self, cuts):
        arr = (C.c_uint32 * (self.world + 1))(*cuts)
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_config_cuts")(self.ctx.h, self.rank, self.world, arr))

    def exchange(self):
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_exchange")(self.ctx.h))

    def celldiv(self):
        info = capi.DivInfo()
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_celldiv")(self.ctx.h, C.byref(info)))
        self.ctx.divinfo = info
        return info

    def forces(self, interstep):
        r = capi.ForceRes()
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_forces")(self.ctx.h, interstep, C.byref(r)))
        return r

    def step(self):
        dt = C.c_double(); r = capi.ForceRes()
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_step")(self.ctx.h, C.byref(dt), C.byref(r)))
        return dt.value, r

    def rebalance(self, bound_weight=0.05, minwidth=None, rank_cost=None):
        ch = C.c_int(0); out = (C.c_uint32 * (self.world + 1))()
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_rebalance")(self.ctx.h, float(bound_weight), float(-1.0 if rank_cost is None else rank_cost), C.byref(ch), out))
        self.cuts = [int(v) for v in out]
        return bool(ch.value)

    def measured_cost(self):
        return float(self.ctx.lib.fn("mgpu_measured_cost")(self.ctx.h))

    def profile(self, on):
        """Switch the per-phase host timers on/off; returns (phases_ms[6], sent_records) accumulated since the last call."""
        ph = (C.c_double * 6)(); sent = C.c_uint64()
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_profile")(self.ctx.h, 1 if on else 0, ph, C.byref(sent)))
        return list(ph), int(sent.value)

    @property
    def sent_records(self):
        sent = C.c_uint64()
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_sent_records")(self.ctx.h, C.byref(sent)))
        return int(sent.value)

    def transport(self):
        """("peer_memory" | "nccl_sendrecv" | "unset", exchanges that fell back to ncclSend/ncclRecv)."""
        d = C.c_int(); fb = C.c_uint64()
        self.ctx.lib.check(self.ctx.lib.fn("mgpu_transport")(self.ctx.h, C.byref(d), C.byref(fb)))
        return {1: "peer_memory", 0: "nccl_sendrecv"}.get(int(d.value), "unset"), int(fb.value)

    def update_divinfo(self):
        """DivInfo of the last divide (dsph_mgpu_step runs its own divides)."""
        info = capi.DivInfo()
        self.ctx.lib.check(self.ctx.lib.fn("celldiv_last_info")(self.ctx.h, C.byref(info)))
        self.ctx.divinfo = info
        return info

    def owned(self):
        st = self.ctx.download()
        keep = (st["code"] & capi.CODE_MASKSPECIAL) == capi.CODE_NORMAL
        return {k: v[keep] for k, v in st.items()}


def slab_select(pos_x: np.ndarray, cuts, rank: int, scell: float, domposmin_x: float):
    """Boolean mask of the particles whose cell column lies in the slab of `rank` (same arithmetic as the cell hash)."""
    cx = ((pos_x - domposmin_x) / scell).astype(np.int64)
    return (cx >= cuts[rank]) & (cx < cuts[rank + 1])
