#!/usr/bin/env python
"""Generates tests/golden/val_dambreak_1s.npz: 1 s of physical time of the 20 k-particle synthetic dam break simulated by
the REFERENCE's stock CPU solver (oracle/_ref/DualSPHysics5.2CPU_linux64, built from /root/reference by oracle/Makefile),
measured every 0.05 s with tests/measure.py (surge front, upstream column height, four pressure gauges).

north_star: "Gauge pressures and surge-front position over 1 s of physical time must be within 1 %".
Run only where /root/reference exists:   python tests/golden/make_validation.py      (about 6 minutes on 4 cores)
"""
import glob
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dualsphysics_multilayer_b200 as dsph  # noqa: E402
import measure  # noqa: E402

REFBIN = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.2CPU_linux64")
NP_TARGET = 20000


def gauges_for(cd):
    ymid = 0.5 * cd.caseposmax[1]
    # G1, G2: bed under the collapsing column; G3: bed downstream (thin tongue); G4: end wall (impact)
    return [(0.1, ymid, 0.03), (0.3, ymid, 0.03), (0.8, ymid, 0.03), (cd.caseposmax[0] - 0.03, ymid, 0.05)]


def series_from_parts(dirout, cd):
    rows = []
    for f in sorted(glob.glob(os.path.join(dirout, "Part_*.bi4"))):
        p, t, step = dsph.load_part(f)
        m = measure.measure_state(p.pos, p.velrhop[:, 3], p.idp, cd.cfg, gauges_for(cd))
        rows.append((t, step, p.np_, m["front"], m["height"], *m["press"]))
    return np.array(rows, np.float64)


if __name__ == "__main__":
    if not os.path.exists(REFBIN):
        sys.exit("oracle/_ref/DualSPHysics5.2CPU_linux64 is missing: run `make -C oracle v52` where /root/reference is mounted")
    tmp = tempfile.mkdtemp(prefix="val_")
    try:
        cd = dsph.make_case("dambreak", NP_TARGET)
        dsph.save_case(cd, tmp, "Case")
        cd = dsph.load_case(tmp, "Case")
        threads = str(min(os.cpu_count() or 1, 8))
        subprocess.check_call([REFBIN, "Case", "out", "-cpu", "-ompthreads:" + threads, "-sv:binx", "-svres:1"], cwd=tmp,
                              stdout=subprocess.DEVNULL)
        s = series_from_parts(os.path.join(tmp, "out"), cd)
        out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "val_dambreak_1s.npz")
        np.savez_compressed(out, series=s, columns=np.array(["time", "step", "np", "front", "height", "g1", "g2", "g3", "g4"]),
                            config=np.frombuffer(bytes(cd.cfg), np.uint8).copy(), np_target=np.uint32(NP_TARGET),
                            gauges=np.array(gauges_for(cd)))
        print(s[:, [0, 1, 3, 4, 5, 6]])
        print("written", out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
