#!/usr/bin/env python
"""torchrun --nproc-per-node N tests/mgpu_gpu_check.py [np] [steps] [kind]: N x-slabs on N GPUs (NCCL) must reproduce the
single-GPU run: identical particle set, same dt, positions/velocities within single-step tolerance."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import mgpu

def main():
    npart = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind = sys.argv[3] if len(sys.argv) > 3 else "dambreak"     # dambreak | nn2layer (multiphase, FDA) [":key=value,..."]
    rebalance = kind.endswith("@rebalance")          # start from bad cuts and let SlabDriver.rebalance() walk them back
    base, _, opts = kind.replace("@rebalance", "").partition(":")
    cd = dsph.make_case(base, npart)
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        setattr(cd.cfg, k, type(getattr(cd.cfg, k))(float(v)))
    code = cd.codes()
    ctx = dsph.create_context(cd, np_capacity=cd.np_ + 200000, device=local)
    scell, dmin, dcells = ctx.domain_info()
    cx = ((cd.pos[:, 0] - dmin[0]) / scell).astype(np.int64)
    cuts = mgpu.balanced_cuts(np.bincount(cx, minlength=dcells[0]), world, 3)
    if rebalance:
        cuts = [cuts[0]] + [min(c + 4, dcells[0] - 3 * (world - i)) for i, c in enumerate(cuts[1:-1], 1)] + [cuts[-1]]
    sel = mgpu.slab_select(cd.pos[:, 0], cuts, rank, scell, dmin[0])
    ctx.upload(cd.idp[sel], cd.pos[sel], cd.velrhop[sel], int(np.count_nonzero(sel[: cd.npb])), code[sel])
    native = os.environ.get("DSPH_MGPU_DRIVER", "native") == "native"      # native = NCCL inside libdsph_b200.so (dsph_mgpu_step)
    drv = mgpu.NativeSlabDriver(ctx, rank, world, cuts) if native else mgpu.SlabDriver(ctx, rank, world, cuts, "cuda:%d" % local)
    drv.celldiv()
    t = 0.0
    moved = 0
    for s in range(nsteps):
        if rebalance and s % 3 == 2 and drv.rebalance(bound_weight=0.3):
            moved += 1
            drv.celldiv()
        dt, _ = drv.step()
        t += dt
    if rebalance and moved == 0:
        print("MGPU_CHECK FAIL: the cuts did not move")
        sys.exit(1)
    own = drv.owned()
    gathered = [None] * world
    transport = drv.transport() if native else ("python", 0)
    dist.all_gather_object(gathered, dict(t=t, idp=own["idp"], pos=own["pos"], velrhop=own["velrhop"], sent=drv.sent_records, transport=transport))
    ok = True
    if rank == 0:
        ref = dsph.create_context(cd, device=local)
        ref.upload(cd.idp, cd.pos, cd.velrhop, cd.npb, code)
        ref.celldiv(True)
        tr = 0.0
        for _ in range(nsteps):
            dt, _ = ref.step()
            tr += dt
        st = ref.download()
        keep = (st["code"] & dsph.capi.CODE_MASKSPECIAL) == dsph.capi.CODE_NORMAL      # periodic copies of the single-GPU run are not particles
        st = {k: v[keep] for k, v in st.items()}
        idp = np.concatenate([g["idp"] for g in gathered]); pos = np.concatenate([g["pos"] for g in gathered]); vr = np.concatenate([g["velrhop"] for g in gathered])
        o1 = np.argsort(idp); o2 = np.argsort(st["idp"])
        same_set = len(idp) == len(st["idp"]) and np.array_equal(idp[o1], st["idp"][o2])
        dpos = float(np.abs(pos[o1] - st["pos"][o2]).max()) if same_set else -1
        dvr = float(np.linalg.norm((vr[o1] - st["velrhop"][o2]).astype(np.float64)) / np.linalg.norm(st["velrhop"][o2].astype(np.float64))) if same_set else -1
        dts = [abs(g["t"] - tr) / tr for g in gathered]
        print("MGPU_CHECK world=%d np=%d steps=%d same_set=%s max|dpos|=%.3e relL2(velrhop)=%.3e dt_rel=%s sent=%s cuts=%s" % (world, cd.np_, nsteps, same_set, dpos, dvr, ["%.1e" % d for d in dts], [g["sent"] for g in gathered], cuts))
        ok = same_set and dpos < 1e-7 and dvr < 1e-5 and max(dts) < 1e-6
        print("MGPU_TRANSPORT", [g["transport"] for g in gathered])
        want = os.environ.get("DSPH_MGPU_EXPECT")                      # "peer_memory" | "nccl_sendrecv" | "fallback"
        if want == "fallback":
            ok = ok and all(g["transport"][0] == "peer_memory" and g["transport"][1] > 0 for g in gathered)
        elif want:
            ok = ok and all(g["transport"][0] == want and g["transport"][1] == 0 for g in gathered)
        print("MGPU_CHECK", "PASS" if ok else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)

if __name__ == "__main__":
    main()
