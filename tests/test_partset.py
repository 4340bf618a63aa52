"""PART output set, host only: Part_Head.ibi4, Part_NNNN.bi4 / Part_pXX_NNNN.bi4 pieces, PartInfo.ibi4 and
PartOut_000.obi4 as written by the driver's PartSet (csrc/host/partset.cpp), checked
  * by an independent Python reader of the JBinaryData container (format: S/JBinaryData.cpp:748-899,1156-1166),
  * by this library's own loader (round trip), and
  * by the REFERENCE CPU solver, which must restart with -partbegin from the files (S/JPartsLoad4.cpp:151-254, S/JSph.cpp:2005-2040)."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import dualsphysics_multilayer_b200 as dsph

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REFCPU = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.2CPU_linux64")

TYPES = {2: ("<i", 4), 3: ("<b", 1), 4: ("<B", 1), 5: ("<h", 2), 6: ("<H", 2), 7: ("<i", 4), 8: ("<I", 4), 9: ("<q", 8), 10: ("<Q", 8),
         11: ("<f", 4), 12: ("<d", 8), 20: ("<3i", 12), 21: ("<3I", 12), 22: ("<3f", 12), 23: ("<3d", 24)}
NPTYPES = {4: np.uint8, 8: np.uint32, 11: np.float32, 12: np.float64, 22: np.float32, 23: np.float64}


class Reader:
    def __init__(self, b):
        self.b, self.c = b, 0

    def u32(self):
        v = struct.unpack_from("<I", self.b, self.c)[0]; self.c += 4; return v

    def i32(self):
        v = struct.unpack_from("<i", self.b, self.c)[0]; self.c += 4; return v

    def s(self):
        n = self.u32(); v = self.b[self.c:self.c + n].decode("latin1"); self.c += n; return v

    def item(self):
        self.u32()
        assert self.s() == "\nITEM\n"
        it = {"name": self.s(), "hide": self.i32(), "hidevalues": self.i32(), "fmtf": self.s(), "fmtd": self.s(), "values": {}, "arrays": {}, "items": []}
        narr, nit, szval = self.u32(), self.u32(), self.u32()
        if szval:
            c0 = self.c
            assert self.s() == "\nVALUES"
            for _ in range(self.u32()):
                name, t = self.s(), self.i32()
                if t == 1:
                    it["values"][name] = self.s()
                else:
                    fmt, sz = TYPES[t]
                    v = struct.unpack_from(fmt, self.b, self.c); self.c += sz
                    it["values"][name] = v[0] if len(v) == 1 else v
            assert self.c - c0 == szval
        for _ in range(narr):
            self.u32()
            assert self.s() == "\nARRAY"
            name, hide, t, cnt, sz = self.s(), self.i32(), self.i32(), self.u32(), self.u32()
            per = 3 if t in (22, 23) else 1
            it["arrays"][name] = np.frombuffer(self.b, NPTYPES[t], cnt * per, self.c).reshape((cnt, 3) if per == 3 else (cnt,))
            self.c += sz
        for _ in range(nit):
            it["items"].append(self.item())
        return it


def read_bi4(path, filecode):
    b = open(path, "rb").read()
    head = b[:64]
    assert head.startswith(("#FileJBD " + filecode).encode()) and head[58:59] == b"\n" and head[60] == 0, head
    r = Reader(b); r.c = 64
    items = []
    while r.c < len(b):
        items.append(r.item())
    return items


@pytest.fixture(scope="module")
def case():
    return dsph.make_case("dambreak", 3000)


def test_partset_files_and_round_trip(tmp_path, case):
    cd = case
    d = str(tmp_path / "parts")
    nout = 5
    out = (np.arange(cd.np_ - nout, cd.np_, dtype=np.uint32) + 100000, np.full((nout, 3), 9.5), np.full((nout, 4), 2.0, np.float32), np.array([1, 2, 3, 1, 2], np.uint8))
    dsph.save_partset(cd, d, 3, 0.0123, step=77, symplecticdtpre=1.5e-4, out=out)
    dsph.save_partset(cd, d, 4, 0.0223, step=99, symplecticdtpre=1.4e-4)
    assert sorted(os.listdir(d)) == ["PartInfo.ibi4", "PartOut_000.obi4", "Part_0003.bi4", "Part_0004.bi4", "Part_Head.ibi4"]
    # Part_Head.ibi4
    (h,) = read_bi4(os.path.join(d, "Part_Head.ibi4"), "JPartDataHead")
    v = h["values"]
    assert v["FmtVersion"] == 180324 and v["Npiece"] == 1 and v["CaseNp"] == cd.cfg.casenp and v["CaseNfluid"] == cd.cfg.casenfluid
    assert v["Dp"] == cd.cfg.dp and abs(v["Gravity"][2] + 9.81) < 1e-6 and v["ViscoType"] == 1
    mk = h["items"][0]
    assert mk["name"] == "MkBlocks" and mk["values"]["Count"] == 2 and [i["values"]["Type"] for i in mk["items"]] == ["Fixed", "Fluid"]
    assert sum(i["values"]["Count"] for i in mk["items"]) == cd.np_
    # Part_0003.bi4
    (p,) = read_bi4(os.path.join(d, "Part_0003.bi4"), "JPartDataBi4")
    assert p["values"]["Npiece"] == 1 and p["values"]["AxisDiv"] == 0
    part = p["items"][0]
    assert part["name"] == "PART_0003" and part["values"]["Npok"] == cd.np_ and part["values"]["TimeStep"] == 0.0123 and part["values"]["Step"] == 77
    assert part["values"]["SymplecticDtPre"] == 1.5e-4 and part["values"]["Nout"] == nout
    assert np.array_equal(part["arrays"]["Idp"], cd.idp) and np.array_equal(part["arrays"]["Posd"], cd.pos)
    assert np.array_equal(part["arrays"]["Vel"], cd.velrhop[:, :3]) and np.array_equal(part["arrays"]["Rhop"], cd.velrhop[:, 3])
    # PartInfo.ibi4: parent (no items) + one PART item without arrays per save
    inf = read_bi4(os.path.join(d, "PartInfo.ibi4"), "JPartDataBi4_Info")
    assert [i["name"] for i in inf] == ["JPartDataBi4", "PART_0003", "PART_0004"]
    assert not inf[0]["items"] and not inf[1]["arrays"] and inf[2]["values"]["TimeStep"] == 0.0223 and "npf" in inf[1]["values"]
    # PartOut_000.obi4: Data record + the excluded particles of PART 3 (PART 4 had none)
    po = read_bi4(os.path.join(d, "PartOut_000.obi4"), "JPartOutBi4")
    assert [i["name"] for i in po] == ["JPartOutBi4", "PART_0003"] and po[0]["values"]["FmtVersion"] == 180320 and po[0]["values"]["Block"] == 0
    assert po[1]["values"]["Nout"] == nout and np.array_equal(po[1]["arrays"]["Motive"], out[3]) and np.array_equal(po[1]["arrays"]["Idp"], out[0])
    # own loader
    c2, t, st, npc = dsph.load_partset(d, 3)
    assert (t, st, npc) == (0.0123, 77, 1) and np.array_equal(c2.pos, cd.pos) and np.array_equal(c2.velrhop, cd.velrhop) and np.array_equal(c2.idp, cd.idp)


def test_partset_pieces(tmp_path, case):
    """Count-split pieces (what a single GPU writes when one file would pass the 4 GB limit of the container)."""
    cd = case
    d = str(tmp_path / "parts")
    piecemax = (cd.np_ + 2) // 3
    dsph.save_partset(cd, d, 1, 0.5, step=10, piecemax=piecemax)
    names = sorted(os.listdir(d))
    assert names == ["PartInfo_p00.ibi4", "PartInfo_p01.ibi4", "PartInfo_p02.ibi4", "PartOut_000.obi4", "Part_Head.ibi4",
                     "Part_p00_0001.bi4", "Part_p01_0001.bi4", "Part_p02_0001.bi4"], names
    tot = 0
    for pc in range(3):
        (p,) = read_bi4(os.path.join(d, "Part_p%02d_0001.bi4" % pc), "JPartDataBi4")
        assert p["values"]["Piece"] == pc and p["values"]["Npiece"] == 3
        tot += p["items"][0]["values"]["Npok"]
    assert tot == cd.np_
    (h,) = read_bi4(os.path.join(d, "Part_Head.ibi4"), "JPartDataHead")
    assert h["values"]["Npiece"] == 3
    c2, t, st, npc = dsph.load_partset(d, 1)
    assert npc == 3 and np.array_equal(c2.pos, cd.pos) and np.array_equal(c2.idp, cd.idp) and np.array_equal(c2.velrhop, cd.velrhop)


def _ref_restart(tmp_path, cd, piecemax):
    dsph.save_case(cd, str(tmp_path), "Case")
    # a state that differs from Case.bi4, so that the restart provably comes from the PART set
    moved = dsph.CaseData(cd.cfg, cd.idp, cd.pos.copy(), cd.velrhop.copy(), cd.npb, cd.name, cd.caseposmin, cd.caseposmax, cd.fluidblocks)
    moved.velrhop[cd.npb:, 0] = 0.25
    dsph.save_partset(moved, str(tmp_path / "parts"), 2, 0.0375, step=40, symplecticdtpre=1.25e-4, piecemax=piecemax)
    r = subprocess.run([REFCPU, "Case", "out", "-cpu", "-ompthreads:2", "-svres:1", "-nsteps:2", "-partbegin:2", "parts"], cwd=str(tmp_path),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "Finished execution (code=0)" in r.stdout, r.stdout[-3000:]
    return r.stdout


@pytest.mark.parametrize("pieces", [1, 2])
def test_reference_cpu_solver_restarts_from_our_partset(tmp_path, pieces):
    if not os.path.exists(REFCPU):
        pytest.skip("reference CPU solver not built here")
    cd = dsph.make_case("dambreak", 3000)
    if pieces > 1:
        # JPartsLoad4::LoadParticles opens piece 0 as "piece 0 of 2" (S/JPartsLoad4.cpp:169) and sizes its arrays as
        # Npiece x Npok(piece 0) (:201-206): the reference itself only restarts from TWO pieces of EQUAL size, so the case is
        # trimmed to an even number of particles (this library's own loader takes any number of pieces of any size)
        n = cd.np_ - cd.np_ % pieces
        drop = cd.np_ - n
        cd = dsph.CaseData(cd.cfg, cd.idp[:n], cd.pos[:n], cd.velrhop[:n], cd.npb, cd.name, cd.caseposmin, cd.caseposmax)
        cd.cfg.casenp -= drop; cd.cfg.casenfluid -= drop
        cd.fluidblocks = [int(cd.cfg.casenfluid)]
    out = _ref_restart(tmp_path, cd, 0 if pieces == 1 else cd.np_ // pieces)
    m = re.search(r"Particles of simulation \(initial\): (\d+)", out)
    assert m and int(m.group(1)) == cd.np_, out[-3000:]
    assert re.search(r"PartBegin=2", out) or "-partbegin" in out or "Part_0002" in out or "0.0375" in out, out[-3000:]
    # the run continues from the PART's time (Part_0003 follows at TimeStep > 0.0375), with the restart velocity in the field
    t = [float(x) for x in re.findall(r"^Part_\d+\s+([0-9.eE+-]+)\s", out, flags=re.M)]
    assert t and min(t) >= 0.0375, (t, out[-3000:])
