#!/usr/bin/env python
"""Motion-law fixtures: what the reference's JDsMotion / JMotion computed at EVERY step of a short run, for <motion> blocks that the
physics goldens (mv_*) do not reach: accelerated rectilinear and rotational movements, chains, a delayed start, two concurrent
movements of one object (translation composed with rotation), angles given in radians, an event with a finish time.

mkcase writes the dam break with the moving plate, the <motion> block of Case.xml is replaced, oracle/_ref/ref_harness52 (the
reference's JSphCpuSingle) runs NSTEPS steps and logs the StMotionData of every step (oracle/ref_harness.cpp: motionlog).
Output: tests/golden/motionlog_<name>.npz {xml, t_dt, val, nobj, casenfixed, casenmoving}.  Needs /root/reference (harness build)."""
import os, re, shutil, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from make_golden import MKCASE, HARNESS, read_dump   # noqa: E402
import dualsphysics_multilayer_b200 as dsph          # noqa: E402

NSTEPS = 40
AX = '<axisp1 x="0.44" y="0" z="0.04" /><axisp2 x="0.44" y="1" z="0.04" />'
CASES = {
    # delayed start, accelerated translation handing its velocity to nothing, then accelerated rotation, then constant rotation
    "chain_rectace_rotace_rot": """<objreal ref="0"><begin mov="1" start="0.0012" />
        <mvrectace id="1" duration="0.003" next="2"><ace x="0" y="0" z="30" /><velini x="0.2" y="0" z="0" /></mvrectace>
        <mvrotace id="2" duration="0.004" next="3">%s<ace ang="4000" /><velini ang="50" /></mvrotace>
        <mvrot id="3" duration="100">%s<vel ang="300" /></mvrot></objreal>""" % (AX, AX),
    # two movements of the same object active at once: the second one starts later and ends by its event's finish time
    "concurrent_rect_rot": """<objreal ref="0"><begin mov="1" start="0" /><begin mov="3" start="0.004" finish="0.011" />
        <mvrect id="1" duration="100"><vel x="0.3" y="0" z="0.1" /></mvrect>
        <mvrot id="3" duration="100">%s<vel ang="200" /></mvrot></objreal>""" % AX,
    # sinusoidal rotation with the angles in radians, previous phase kept by the second link of the chain
    "rotsinu_radians_chain": """<objreal ref="0"><begin mov="1" start="0" />
        <mvrotsinu id="1" duration="0.006" next="2" anglesunits="radians">%s<freq v="20" /><ampl v="0.2" /><phase v="0.3" /></mvrotsinu>
        <mvrotsinu id="2" duration="100" anglesunits="radians">%s<freq v="35" /><ampl v="0.1" /></mvrotsinu></objreal>""" % (AX, AX),
    # wait in the middle of a chain, accelerated translation that keeps the velocity of the previous link
    "rect_wait_rectace_prev": """<objreal ref="0"><begin mov="1" start="0" />
        <mvrectace id="1" duration="0.003" next="2"><ace x="40" y="0" z="0" /><velini x="0" y="0" z="0" /></mvrectace>
        <wait id="2" duration="0.002" next="3" />
        <mvrectsinu id="3" duration="100"><freq x="30" y="0" z="12" /><ampl x="0.01" y="0" z="0.02" /><phase x="20" y="0" z="70" /></mvrectsinu></objreal>""",
    # circular movements: the object is translated along a circle (no rotation), constant / accelerated / sinusoidal angle
    "chain_cir_cirace_cirsinu": """<objreal ref="0"><begin mov="1" start="0.0005" />
        <mvcir id="1" duration="0.004" next="2">%s<ref x="0.44" y="0.3" z="0.2" /><vel ang="400" /></mvcir>
        <mvcirace id="2" duration="0.004" next="3">%s<ref x="0.44" y="0.3" z="0.2" /><ace ang="9000" /></mvcirace>
        <mvcirsinu id="3" duration="100">%s<ref x="0.44" y="0.3" z="0.2" /><freq v="40" /><ampl v="8" /><phase v="30" /></mvcirsinu></objreal>""" % (AX, AX, AX),
    # file-driven movements (wave-maker style): positions from a 4-column file with remarks and mixed separators, angles in radians
    "rectfile_then_rotfile": """<objreal ref="0"><begin mov="1" start="0.001" />
        <mvrectfile id="1" duration="0.006" next="2"><file name="piston.dat" fields="4" fieldtime="0" fieldx="1" fieldz="3" /></mvrectfile>
        <mvrotfile id="2" duration="100" anglesunits="radians"><file name="flap.dat" />%s</mvrotfile></objreal>""" % AX,
    # nested objects: a group that translates (and later also rotates) carries a real object that rotates about its own axis; the
    # child's axis travels with the parent
    "nested_parent_rect_child_rot": """<obj><begin mov="1" start="0" /><begin mov="5" start="0.008" />
        <mvrect id="1" duration="100"><vel x="0.5" y="0" z="0.2" /></mvrect>
        <mvrot id="5" duration="100"><axisp1 x="0.2" y="0" z="0.5" /><axisp2 x="0.2" y="1" z="0.5" /><vel ang="150" /></mvrot>
        <objreal ref="0"><begin mov="1" start="0.002" />
          <mvrot id="1" duration="0.009" next="2">%s<vel ang="300" /></mvrot>
          <mvcir id="2" duration="100">%s<ref x="0.44" y="0.3" z="0.2" /><vel ang="250" /></mvcir></objreal></obj>""" % (AX, AX),
}

DATAFILES = {
    "piston.dat": "# time; x; unused; z\n" + "".join("%.6f;%.8f;7;%.8f\n" % (0.002 + 0.0007 * i, 0.004 * (i ** 1.3) * 0.1, 0.002 * i * (1 if i % 2 else -0.5)) for i in range(14)),
    "flap.dat": "# time angle[rad]\n" + "".join("%.6f\t%.8f\n" % (0.0009 * i, 0.05 * ((i * 0.37) % 1.0 - 0.3)) for i in range(30)),
}


def run(name, body, outdir):
    tmp = tempfile.mkdtemp(prefix="motionlog_")
    try:
        subprocess.check_call([MKCASE, "movgate", "3000", tmp, "Case", "tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2"], stdout=subprocess.DEVNULL)
        xmlfile = os.path.join(tmp, "Case.xml")
        xml = open(xmlfile).read()
        xml, n = re.subn(r"<motion>.*</motion>", "<motion>\n" + body + "\n    </motion>", xml, flags=re.S)
        assert n == 1
        open(xmlfile, "w").write(xml)
        used = {}
        for fn, text in DATAFILES.items():
            if fn in body:
                open(os.path.join(tmp, fn), "w").write(text); used[fn] = text
        cmd = [HARNESS, "Case", "out", "-cpu", "-ompthreads:1", "-sv:none", "-hdir:" + os.path.join(tmp, "dump"), "-hsteps:%d" % (NSTEPS + 5), "-hnsteps:%d" % NSTEPS]
        r = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout[-3000:]); raise RuntimeError("reference harness failed for " + name)
        d = read_dump(os.path.join(tmp, "dump", "s%04d_motionlog.dmp" % NSTEPS))
        cd = dsph.load_case(tmp, "Case")
        out = os.path.join(outdir, "motionlog_" + name + ".npz")
        np.savez_compressed(out, xml=np.frombuffer(xml.encode(), np.uint8), t_dt=d["t_dt"], val=d["val"], nobj=d["nobj"],
                            casenfixed=np.uint32(cd.cfg.casenfixed), casenmoving=np.uint32(cd.cfg.casenmoving),
                            **{"file_" + fn: np.frombuffer(text.encode(), np.uint8) for fn, text in used.items()})
        val = np.asarray(d["val"]).reshape(-1, 23)
        print("%-28s %d steps, types %s, %.1f KB" % (name, val.shape[0], sorted(set(int(v) for v in val[:, 0])), os.path.getsize(out) / 1024))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    for k, v in CASES.items():
        if len(sys.argv) < 2 or k in sys.argv[1:]:
            run(k, v, out)
