#!/usr/bin/env python
"""bench.py -- throughput of the per-step particle hot path (cell list -> Interaction_Forces -> ComputeStep).

    python bench.py --gpus N --steps K --warmup W [--np NP] [--impl reference]

One "step" = one full Symplectic time step (2 force evaluations, 2 ComputeStep updates, 2 neighbour-list builds) of
the synthetic 3-D dam-break DB(N) of SURVEY.md 8(d) (BASELINE.json configs[1]: 20 M particles on 1 B200).
Prints ONE JSON line (contract in the task statement): value = M particle-steps/s with the state resident in HBM,
timed with CUDA events on the solver's stream; e2e = the same metric through the C ABI with HOST buffers (pinned H2D
of the particle state + D2H of the result inside the timed region); roofline = gather-byte model of the fluid force
kernel against the measured HBM copy bandwidth; cpu_baseline = the reference's own OpenMP CPU solver
(oracle/_ref) on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Mparticle-steps/s (3D dam-break)"
UNIT = "Mparticle-steps/s"
REFBIN = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.2CPU_linux64")
MKCASE = os.path.join(ROOT, "dualsphysics_multilayer_b200", "bin", "mkcase")


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.sm, self.maxsm, self.reasons = gpu_index, False, [], 0, set()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                if len(f) >= 6:
                    self.sm.append(float(f[0])); self.maxsm = float(f[1])
                    for n, v in zip(names, f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        sm = sorted(self.sm)
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.maxsm or None, "reasons": sorted(self.reasons)}


WORKLOADS = {
    "dambreak": ("3D dam-break", "synthetic 3D dam-break DB(N) (SURVEY.md 8d): Symplectic + Wendland + artificial viscosity 0.1 + DDT2 0.1, DBC walls"),
    "periodicbox": ("3D periodic box", "synthetic periodic fluid box PB(N) (SURVEY.md 8d): Taylor-Green field, Symplectic + Wendland + artificial viscosity + DDT1, periodic in x, y, z"),
    "nn2layer": ("two-layer non-Newtonian dam-break", "synthetic two-layer multiphase dam-break NN(N) (SURVEY.md 8d): FDA velocity gradients, Morris operator, DDT3, shifting"),
}


def run_reference(np_sample, nsteps, threads, kind="dambreak"):
    """Times the reference's own CPU implementation (JSphCpuSingle, OpenMP) on the workload at np_sample for nsteps steps."""
    refbin = REFBIN if kind != "nn2layer" else os.path.join(os.path.dirname(REFBIN), "DualSPHysics5.0_NNewtonianCPU_linux64")
    if not os.path.exists(refbin):
        return None
    tmp = tempfile.mkdtemp(prefix="bench_ref_")
    try:
        subprocess.check_call([MKCASE, kind, str(np_sample), tmp, "Case"], stdout=subprocess.DEVNULL)
        r = subprocess.run([refbin, "Case", "out", "-cpu", "-ompthreads:%d" % threads, "-sv:none", "-svtimers:1",
                            "-nsteps:%d" % nsteps], cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        txt = r.stdout
        m1 = re.search(r"Simulation Runtime\.*: ([0-9.eE+-]+) sec", txt)
        m2 = re.search(r"Steps of simulation\.*: (\d+)", txt)
        m3 = re.search(r"Particles of simulation \(initial\): (\d+)", txt)
        if not (m1 and m2 and m3):
            return None
        t, steps, npart = float(m1.group(1)), int(m2.group(1)), int(m3.group(1))
        forces = re.search(r"CF-Forces\.*: ([0-9.eE+-]+) sec", txt)
        return dict(value=npart * steps / t / 1e6, np=npart, steps=steps, seconds=t,
                    cf_forces_s=(float(forces.group(1)) if forces else None))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


REFGPU = os.path.join(os.path.dirname(REFBIN), "DualSPHysics5.2_linux64")


def reference_gpu_baseline(npart, kind="dambreak", steps_a=8, steps_b=28):
    """The reference's OWN GPU solver (JSphGpuSingle + its CUDA kernels, S/Makefile flags) recompiled for sm_100 by
    `make -C oracle gpu52`, run on this GPU on the same workload: steady-state rate from two runs of different length (their
    difference removes initialisation and first-launch effects).  A reported baseline: what the unmodified reference kernels
    deliver on a B200."""
    if kind == "nn2layer" or not os.path.exists(REFGPU):
        return None
    tmp = tempfile.mkdtemp(prefix="bench_refgpu_")
    try:
        subprocess.check_call([MKCASE, kind, str(npart), tmp, "Case"], stdout=subprocess.DEVNULL)
        res = []
        for ns in (steps_a, steps_b):
            r = subprocess.run([REFGPU, "Case", "out", "-gpu", "-sv:none", "-svtimers:0", "-nsteps:%d" % ns], cwd=tmp,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            m1 = re.search(r"Simulation Runtime\.*: ([0-9.eE+-]+) sec", r.stdout)
            m2 = re.search(r"Steps of simulation\.*: (\d+)", r.stdout)
            m3 = re.search(r"Particles of simulation \(initial\): (\d+)", r.stdout)
            if not (m1 and m2 and m3):
                return {"value": None, "unit": UNIT, "what": "reference GPU solver failed: " + r.stdout[-200:].replace("\n", " ")}
            res.append((float(m1.group(1)), int(m2.group(1)), int(m3.group(1))))
        (ta, na, npa), (tb, nb, _) = res
        rate = npa * (nb - na) / (tb - ta) / 1e6 if tb > ta and nb > na else None
        # one more short run with the reference's synchronising timers on, only for its own split of the step
        r = subprocess.run([REFGPU, "Case", "out", "-gpu", "-sv:none", "-svtimers:1", "-nsteps:%d" % steps_b], cwd=tmp,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        split = {}
        for key in ("NL-Limits", "NL-PreSort", "NL-RadixSort", "NL-CellBegin", "NL-SortData", "CF-Forces", "SU-ComputeStep"):
            m = re.search(re.escape(key) + r"\.*: ([0-9.eE+-]+) sec", r.stdout)
            if m:
                split[key] = 1e3 * float(m.group(1)) / steps_b
        return {"value": rate, "unit": UNIT, "np": npa, "ms_per_step": 1e3 * (tb - ta) / (nb - na) if nb > na else None,
                "its_own_timers_ms_per_step": split,
                "what": "reference DualSPHysics5.2 GPU solver (its own CUDA kernels, nvcc 12.9 -O3 -use_fast_math, sm_100) on this GPU, "
                        "%s at %d particles: steps %d..%d of two runs (%.2f s and %.2f s of Simulation Runtime)"
                        % (WORKLOADS[kind][0], npa, na, nb, ta, tb)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cpu_baseline(sample_seconds=25.0, steps=3, kind="dambreak"):
    cores = os.cpu_count() or 1
    threads = min(cores, 64)                      # OMP_MAXTHREADS 64, S/OmpDefs.h:39
    np_sample = int(min(20_000_000, max(200_000, 0.05e6 * threads * sample_seconds / steps)))
    r = run_reference(np_sample, steps, threads, kind)
    if r is None:
        return {"value": None, "unit": UNIT, "cores": threads, "kind": "reference", "sample": "oracle/_ref binary unavailable"}
    return {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": "reference DualSPHysics CPU solver (OpenMP, %d threads): %s at %d particles, %d steps in %.1f s (CF-Forces %.1f s)"
                      % (threads, WORKLOADS[kind][0], r["np"], r["steps"], r["seconds"], r["cf_forces_s"] or -1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--np", "--particles", dest="np", type=int, default=20_000_000, help="particles per GPU (DB(N))")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pre-steps", type=int, default=0, help="untimed steps before warm-up (to leave the initial lattice)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the run of the reference's own GPU solver (oracle/_ref/DualSPHysics5.2_linux64)")
    ap.add_argument("--force-variant", type=int, default=0)
    ap.add_argument("--workload", default="dambreak", choices=["dambreak", "periodicbox", "nn2layer"],
                    help="DB(N) is the headline (BASELINE.json); PB(N) and NN(N) are the other single-GPU configs of SURVEY.md 8(d)")
    ap.add_argument("--profile-mgpu", action="store_true", help="N>1: 3 extra untimed steps with per-phase host timers (adds syncs)")
    ap.add_argument("--bound-weight", type=float, default=0.05, help="N>1: cost of a boundary particle relative to a fluid one when balancing the slabs")
    ap.add_argument("--balance-steps", type=int, default=3, help="N>1: untimed steps before warm-up, each followed by SlabDriver.rebalance with the measured per-rank kernel time")
    a = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): everything else any library prints (e.g. NCCL's
    # version banner) goes to stderr; the JSON is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return
        cores = os.cpu_count() or 1
        threads = min(cores, 64)
        nsteps = max(1, a.steps + a.warmup)
        np_sample = int(min(a.np, max(200_000, 0.05e6 * threads * 60.0 / nsteps)))
        t0 = time.time()
        r = run_reference(np_sample, nsteps, threads, a.workload)
        if r is None:
            emit(({"impl": "reference", "unavailable": "oracle/_ref/DualSPHysics5.2CPU_linux64 is missing (built by oracle/Makefile where /root/reference is mounted)"}))
            return
        sample = ("reference DualSPHysics5.2 CPU solver (OpenMP, %d threads) on dam-break DB(%d) = bounded sample of the %d-particle workload, %d steps, %.1f s"
                  % (threads, r["np"], a.np, r["steps"], r["seconds"]))
        emit(({"impl": "reference", "metric": METRIC.replace("3D dam-break", WORKLOADS[a.workload][0]), "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": a.warmup, "ms_per_step": 1e3 * r["seconds"] / r["steps"], "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": WORKLOADS[a.workload][1], "np_sample": r["np"], "np_workload": a.np},
                          "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
                          "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "wall_s": time.time() - t0}))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import dualsphysics_multilayer_b200 as dsph

    if world > 1 and a.workload != "dambreak":
        raise SystemExit("bench.py: the x-slab runs use the dam-break workload")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    L = dsph.lib()
    l = L.lib
    l.dsph_timer_start.argtypes = [C.c_void_p]
    l.dsph_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    l.dsph_fluid_kernel_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]

    from dualsphysics_multilayer_b200 import mgpu
    drv = None
    if world == 1:
        cd = dsph.make_case(a.workload, a.np)
        n = cd.np_
        code = cd.codes()
        ctx = dsph.create_context(cd, device=local)
        if a.force_variant:
            l.dsph_set_force_variant(ctx.h, a.force_variant)
        ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, code)
        ctx.celldiv(True)
        stepfn = ctx.step
        nglobal = n
    else:
        # weak scaling: DB(world*np) split into x-slabs of balanced work (fluid 1, boundary w); every rank generates only its slab
        nglobal_target = a.np * world
        info = dsph.make_case_slab("dambreak", nglobal_target, 0, 0)          # constants + counts of the full case
        sc = C.c_double(); pm = (C.c_double * 3)(); dcl = (C.c_uint32 * 3)()
        L.check(L.fn("domain_info")(C.byref(info.cfg), C.byref(sc), pm, dcl))
        ncol = int(dcl[0])
        # particles per lattice x-index of DB(N) (mkcase geometry: walls at index 0 and NX, bottom at k=0, water column
        # a x 1.675a x 0.75a at x=0), mapped to cell columns with the arithmetic of the cell hash
        dp = info.cfg.dp; a_len = 0.4
        NX, NY, NZ = (int(round(v / dp)) for v in info.caseposmax)
        FX, FY, FZ = min(int(round(a_len / dp)), NX - 1), NY - 1, min(int(round(0.75 * a_len / dp)), NZ - 1)
        ii = np.arange(NX + 1)
        percol = a.bound_weight * np.where((ii == 0) | (ii == NX), (NY + 1) * (NZ + 1), (NY + 1) + 2 * NZ).astype(np.float64)
        percol[1:FX + 1] += FY * FZ                                     # work-weighted: fluid 1, boundary --bound-weight
        colidx = ((ii * dp - pm[0]) / sc.value).astype(np.int64)
        colcount = np.bincount(colidx, weights=percol, minlength=ncol)
        cuts = mgpu.balanced_cuts(colcount, world, 4)
        cd = dsph.make_case_slab("dambreak", nglobal_target, cuts[rank], cuts[rank + 1])
        n = cd.np_
        code = cd.codes()
        ctx = dsph.create_context(cd.cfg, int(n * 1.5) + 200_000, local)
        if a.force_variant:
            l.dsph_set_force_variant(ctx.h, a.force_variant)
        ctx.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, code)
        drv = mgpu.SlabDriver(ctx, rank, world, cuts, "cuda:%d" % local)
        drv.celldiv()
        stepfn = drv.step
        nglobal = int(info.cfg.casenp)
    for _ in range(a.pre_steps):
        stepfn()
    cuts_initial = list(drv.cuts) if drv is not None else None
    if drv is not None:
        # dynamic load balance (setup, untimed): the static weights cannot know that free-surface and wall particles have fewer
        # neighbours; move the cuts with the measured device time of every rank's own kernels, as a production run does every k steps
        for _ in range(a.balance_steps):
            drv.measured_cost()
            stepfn()
            if drv.rebalance(bound_weight=a.bound_weight, rank_cost=drv.measured_cost()):
                drv.celldiv()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: W warm-up + K timed steps, CUDA events on the solver's stream
    for _ in range(a.warmup):
        stepfn()
    tm = dsph.Timers()
    l.dsph_get_timers(ctx.h, C.byref(tm), 1)
    l.dsph_fluid_kernel_stats(ctx.h, None, None, 1)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    L.check(l.dsph_timer_start(ctx.h))
    pstep = 0
    for _ in range(a.steps):
        stepfn()
        pstep += (ctx.np_ if drv is None else ctx.np_ - int(ctx.divinfo.npbper) - int(ctx.divinfo.npfper))   # owned particles only
    ms = C.c_double()
    L.check(l.dsph_timer_stop(ctx.h, C.byref(ms)))
    barrier()
    sampler.stop_flag = True; sampler.join(2.0)
    l.dsph_get_timers(ctx.h, C.byref(tm), 0)
    fms = C.c_double(); fnl = C.c_uint64()
    l.dsph_fluid_kernel_stats(ctx.h, C.byref(fms), C.byref(fnl), 0)
    t_ms = torch.tensor([ms.value], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(pstep)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    value = tot.item() / (t_ms.item() * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (fluid Interaction_Forces): gather-byte model, SURVEY.md 8(d)
    if drv is None:
        ctx.celldiv(False)
    fr_real, fr_chk, br_real, br_chk = ctx.pips()
    npf = ctx.np_ - ctx.npb - (0 if drv is None else int(ctx.divinfo.npfper))
    bytes_per_launch = 58.0 * npf + 16.0 * (fr_chk + fr_real) + (2.0 * fr_real if a.workload == "nn2layer" else 0.0)   # + code[p2]: phase of the neighbour
    k_ms = fms.value / max(1, fnl.value)
    peak, peaksrc = measured_peak()
    achieved = bytes_per_launch / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    # DRAM bytes of one launch of this kernel from the committed `ncu --set full` capture of this very workload (profiles/)
    traffic, traffic_src = None, None
    tfile = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01d_forces_v9_20m_summary.json")
    ncu_summary = None
    if world == 1 and a.np == 20_000_000 and not a.force_variant and os.path.exists(tfile):
        ncu_summary = json.load(open(tfile))
        traffic = ncu_summary["dram_bytes_per_launch"]
        traffic_src = "profiles/r01d_forces_v9_20m_summary.json (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel": "ker_forces_fluid", "avg_launch_ms": k_ms, "launches": int(fnl.value),
                "model": "B_gather = 58*Nf + 16*(N_chk+N_real) bytes per launch (SURVEY.md 8d)",
                "pairs_per_fluid_particle": {"checked": fr_chk / max(npf, 1), "real": fr_real / max(npf, 1)},
                "fluid_evals_per_s": npf / (k_ms * 1e-3) if k_ms > 0 else 0.0, "peak_source": peaksrc,
                "share_of_step": fms.value / ms.value if ms.value > 0 else None}
    if ncu_summary is not None:
        # what actually limits the kernel (from the same committed capture): the gather bytes are served by L1/L2, DRAM only sees `traffic`
        m = ncu_summary["metrics"]
        roofline["limiter"] = {"what": "instruction issue", "smsp__issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"].split()[0]),
                               "warp_instructions_per_launch": float(m["smsp__inst_executed.sum"].split()[0]),
                               "threads_per_instruction": float(m["smsp__thread_inst_executed_per_inst_executed.ratio"].split()[0]),
                               "l1_hit_pct": float(m["l1tex__t_sector_hit_rate.pct"].split()[0]), "l2_hit_pct": float(m["lts__t_sector_hit_rate.pct"].split()[0]),
                               "dram_gbs": traffic / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None}

    # ---- end to end through the C ABI with HOST buffers: H2D of the state, one step, D2H of the result, every step
    e2e = None
    if not a.no_e2e:
        st = ctx.download() if drv is None else drv.owned()
        hid = torch.from_numpy(st["idp"]).pin_memory(); hcode = torch.from_numpy(st["code"].view(np.int16)).pin_memory()
        hpos = torch.from_numpy(st["pos"]).pin_memory(); hvr = torch.from_numpy(st["velrhop"]).pin_memory()
        npb = ctx.npb if drv is None else int(np.count_nonzero((st["code"] & 0x1800) != 0x1800))
        nn = int(st["idp"].shape[0])
        cap_out = int(nn * 1.3) + 100_000
        oid = torch.empty(cap_out, dtype=torch.int32).pin_memory(); ocode = torch.empty(cap_out, dtype=torch.int16).pin_memory()
        opos = torch.empty((cap_out, 3), dtype=torch.float64).pin_memory(); ovr = torch.empty((cap_out, 4), dtype=torch.float32).pin_memory()
        up = L.fn("particles_upload"); down = L.fn("particles_download")

        def make_out():
            return (torch.empty(cap_out, dtype=torch.int32).pin_memory(), torch.empty(cap_out, dtype=torch.int16).pin_memory(),
                    torch.empty((cap_out, 3), dtype=torch.float64).pin_memory(), torch.empty((cap_out, 4), dtype=torch.float32).pin_memory())

        compute_lock = threading.Lock()       # kernels of the two pipelines take turns; copies run outside the lock

        def e2e_step(cx, dr, outb):
            oid, ocode, opos, ovr = outb
            L.check(up(cx.h, nn, npb, C.cast(hid.data_ptr(), C.POINTER(C.c_uint32)), C.cast(hcode.data_ptr(), C.POINTER(C.c_uint16)),
                       C.cast(hpos.data_ptr(), C.POINTER(C.c_double)), C.cast(hvr.data_ptr(), C.POINTER(C.c_float))))
            if dr is None:
                with compute_lock:
                    cx.celldiv(True)
                    cx.step()
            else:
                dr.celldiv()
                dr.step()
            m = min(cx.np_, cap_out)
            L.check(down(cx.h, 0, m, C.cast(oid.data_ptr(), C.POINTER(C.c_uint32)), C.cast(ocode.data_ptr(), C.POINTER(C.c_uint16)), None,
                         C.cast(opos.data_ptr(), C.POINTER(C.c_double)), C.cast(ovr.data_ptr(), C.POINTER(C.c_float))))
            return m if dr is None else m - int(cx.divinfo.npbper) - int(cx.divinfo.npfper)

        # Single GPU: two contexts (each with its own CUDA stream and device buffers) driven by two host threads, so that
        # the H2D/D2H copies of one batch overlap the kernels of the other -- what a caller with a stream of independent
        # states does with this API.  Every step still pays its own H2D of the full state and D2H of the full result.
        # Slab runs (N>1) keep one pipeline per rank (the exchange is collective).
        pipes = [(ctx, drv, (oid, ocode, opos, ovr))]
        if drv is None:
            ctx2 = dsph.create_context(cd.cfg, int(n * 1.2) + 100_000, local)
            if a.force_variant:
                l.dsph_set_force_variant(ctx2.h, a.force_variant)
            pipes.append((ctx2, None, make_out()))
        for cx, dr, ob in pipes:
            e2e_step(cx, dr, ob)
        ke = max(2, min(a.steps, 5))
        done = [0] * len(pipes)

        def worker(i):
            cx, dr, ob = pipes[i]
            for _ in range(ke):
                done[i] += e2e_step(cx, dr, ob)
        barrier()
        t0 = time.perf_counter()
        if len(pipes) == 1:
            worker(0)
        else:
            ths = [threading.Thread(target=worker, args=(i,)) for i in range(len(pipes))]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        barrier()
        wall = time.perf_counter() - t0
        pe = sum(done)
        te = torch.tensor([wall], dtype=torch.float64, device="cuda")
        pt = torch.tensor([float(pe)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX); dist.all_reduce(pt, op=dist.ReduceOp.SUM)
        if len(pipes) > 1:
            pipes[1][0].close()
        e2e = {"value": pt.item() / te.item() / 1e6, "unit": UNIT, "h2d_bytes_per_step": nn * (4 + 2 + 24 + 16),
               "d2h_bytes_per_step": nn * (4 + 2 + 24 + 16), "steps": ke * len(pipes),
               "what": "dsph_particles_upload (pinned host -> device) + dsph_celldiv_run + dsph_step + dsph_particles_download per step"
                       + ("; two contexts on two host threads so that the copies of one batch overlap the kernels of the other" if len(pipes) > 1 else ""),
               "timed": "host wall clock around all steps, device synchronised on both sides"}

    phases = None
    if drv is not None and a.profile_mgpu:
        drv.prof = {}
        for _ in range(3):
            drv.step()
        mine = {k: 1e3 * v / 3 for k, v in drv.prof.items()}
        mine["np_owned"] = ctx.np_ - int(ctx.divinfo.npbper) - int(ctx.divinfo.npfper)
        mine["np_ghost"] = int(ctx.divinfo.npbper) + int(ctx.divinfo.npfper)
        phases = [None] * world
        dist.all_gather_object(phases, mine)
        drv.prof = None

    if rank == 0:
        out = {"metric": METRIC.replace("3D dam-break", WORKLOADS[a.workload][0]), "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
               "ms_per_step": t_ms.item() / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": WORKLOADS[a.workload][1],
                          "np_per_gpu": n, "nfluid": int(cd.cfg.casenfluid), "pre_steps": a.pre_steps,
                          "np_global": nglobal,
                          "parallelism": ("single GPU" if world == 1 else "%d x-slabs of balanced work (fluid 1, boundary %g); per divide: NCCL send/recv of migrants + 2h halo copies with <=2 neighbours; per force evaluation: all-reduce(max) of the dt maxima; cuts re-balanced %d times with the measured kernel time per rank before warm-up (%s -> %s)" % (world, a.bound_weight, a.balance_steps, cuts_initial, list(drv.cuts))),
                          "l2": "working set (np*~230 B) is far larger than the 126 MB L2; no flush needed"},
               "roofline": roofline, "e2e": e2e, "gpu_launches": int(tm.kernel_launches), "clocks": sampler.result(),
               "timers_ms_per_step": {k: getattr(tm, k) / a.steps for k in ("nl_limits", "nl_presort", "nl_radixsort", "nl_cellbegin", "nl_sortdata",
                                                                             "cf_preforces", "cf_forces", "su_computestep")}}
        if phases is not None:
            out["mgpu_phases_ms_per_step"] = phases
        if not a.no_cpu_baseline and world == 1:                         # the CPU baseline is reported at N=1 only
            out["cpu_baseline"] = cpu_baseline(kind=a.workload)
            if not a.no_reference_gpu:
                ctx.close()                                              # free the device before the reference solver allocates
                out["reference_gpu"] = reference_gpu_baseline(a.np, a.workload)
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
