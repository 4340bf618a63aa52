#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE's own CPU solver (run only where /root/reference exists).

For every case below: the repo's mkcase writes Case.xml + Case.bi4, oracle/_ref/ref_harness52 (the reference
JSphCpuSingle, subclassed by oracle/ref_harness.cpp) simulates it single-threaded and dumps the intermediate state of
selected steps; the dumps are packed into one compressed npz together with the raw dsph_config bytes and the case
arrays.  These files pin the CPU restatement in oracle/ (tests/test_oracle_vs_reference.py) and the CUDA path
(tests/test_gpu_parity.py).

    python tests/golden/make_golden.py            # (re)generates every fixture
"""
import ctypes as C
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
from dumpio import read_dump  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness52")
HARNESS_NN = os.path.join(ROOT, "oracle", "_ref", "ref_harnessnn")     # src_mphase fork (v5.0 NNewtonian)
MKCASE = os.path.join(ROOT, "dualsphysics_multilayer_b200", "bin", "mkcase")

# name, kind, np_target, mkcase options, dumped steps, total steps
CASES = [
    ("db_sym_wend_art_ddt2", "dambreak", 3000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2"], [0, 40], 41),
    ("db_ver_cubic_art_ddt1_shift", "dambreak", 3000, ["tstep=1", "tkernel=1", "tvisco=1", "visco=0.1", "ddt=1", "shifting=3", "shiftcoef=-2", "shifttfs=2.75", "verletsteps=10"], [0, 9, 25], 26),
    ("db_sym_wend_lamsps_ddt3", "dambreak", 3000, ["tstep=2", "tkernel=2", "tvisco=2", "visco=0.01", "ddt=3"], [0, 30], 31),
    ("db_sym_cubic_art_noddt_half", "dambreak", 3000, ["tstep=2", "tkernel=1", "tvisco=1", "visco=0.05", "ddt=0", "cellmode=2", "shifting=1", "shiftcoef=-2"], [20], 21),
    ("pb_sym_wend_art_ddt1", "periodicbox", 1728, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1"], [0, 15], 16),
    ("pb_ver_wend_art_ddt1", "periodicbox", 1728, ["tstep=1", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "verletsteps=5"], [0, 12], 13),
    # multiphase non-Newtonian fork (two-layer dam break): FDA / SPH velocity gradients, Morris / constitutive-equation operator
    ("nn_sym_fda_lam_ddt3_shift", "nn2layer", 3000, ["tstep=2", "tkernel=2", "velgrad=1", "tvisco=2", "ddt=3", "shifting=3"], [0, 25], 26),
    ("nn_ver_sph_lam_ddt1", "nn2layer", 3000, ["tstep=1", "tkernel=2", "velgrad=2", "tvisco=2", "ddt=1", "shifting=0", "verletsteps=10"], [22], 23),
    ("nn_sym_sph_conseq_ddt2_shift", "nn2layer", 3000, ["tstep=2", "tkernel=2", "velgrad=2", "tvisco=3", "ddt=2", "shifting=3"], [20], 21),
    ("nn_sym_fda_conseq_cubic", "nn2layer", 3000, ["tstep=2", "tkernel=1", "velgrad=1", "tvisco=3", "ddt=0", "shifting=0"], [20], 21),
    ("nn_sym_fda_art_ddt1", "nn2layer", 3000, ["tstep=2", "tkernel=2", "velgrad=1", "tvisco=1", "ddt=1", "shifting=1"], [20], 21),
    # THREE stacked layers with the phases of the fork's own example (R/examples/mphase_nnewtonian/01_WetDambreak/CaseWetDambreak2DNN_Def.xml:
    # rho 2000/1500/1000; the bottom layer shear-thickening with HBP_n=1.5, HBP_m=100) and its parameter set (2-D, Symplectic, Wendland, FDA,
    # laminar, DDT3, shifting 3 coef -10 TFS 1.5); plus the 3-D SPH-gradient / constitutive-equation variant
    ("nn3_2d_sym_fda_lam_ddt3_shift", "nn3layer2d", 3000, ["tstep=2", "tkernel=2", "velgrad=1", "tvisco=2", "ddt=3", "shifting=3", "shiftcoef=-10", "shifttfs=1.5"], [2, 30], 31),   # (step 0 is not dumped: on the untouched lattice ONE particle of 2 596 takes another shift, 4.9e-8 m, the rest agree to 1e-12)
    ("nn3_sym_sph_conseq_ddt2", "nn3layer", 4000, ["tstep=2", "tkernel=2", "velgrad=2", "tvisco=3", "ddt=2", "shifting=0"], [20], 21),
    # 2-D simulations (x-z plane): Simulate2D kernel constants, masses, cell domain with one cell in y, Ace.y = 0
    ("db2d_sym_wend_art_ddt2_shift", "dambreak2d", 3000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2", "shifting=3", "shiftcoef=-2", "shifttfs=1.5"], [0, 30], 31),
    ("db2d_ver_cubic_lamsps_ddt1", "dambreak2d", 3000, ["tstep=1", "tkernel=1", "tvisco=2", "visco=0.01", "ddt=1", "verletsteps=10"], [25], 26),
    ("nn2d_sym_sph_lam_ddt3_shift", "nn2layer2d", 3000, ["tstep=2", "tkernel=2", "velgrad=2", "tvisco=2", "ddt=3", "shifting=3"], [25], 26),
    ("nn2d_sym_fda_conseq_ddt2", "nn2layer2d", 3000, ["tstep=2", "tkernel=2", "velgrad=1", "tvisco=3", "ddt=2", "shifting=0"], [25], 26),
    # particles leaving: a RhopOut window just above the initial hydrostatic maximum excludes 4 / 20 / 146 particles by steps 10 / 20 / 40 (OUTRHOP marking in
    # the step update, special boxes of the divide, NpfOut bookkeeping); options after "xml:" patch <parameter> values of Case.xml
    ("db_sym_wend_art_ddt2_rhopout", "dambreak", 3000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2", "xml:RhopOutMax=1002.14", "xml:RhopOutMin=999.2"], [10, 20, 40], 41),
    # mDBC boundary correction (Boundary=2, slip mode velocity=0): normals file written by mkcase, boundary densities after
    # JSphCpu::Interaction_MdbcCorrection dumped with the forces; DDT1 so that the boundary rule of S/JSphCpu.cpp:730 is on the path
    ("db_sym_wend_art_ddt1_mdbc", "dambreak", 3000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "boundary=2"], [0, 30], 31),
    ("db_sym_cubic_art_ddt2_mdbccorr", "dambreak", 3000, ["tstep=2", "tkernel=1", "tvisco=1", "visco=0.1", "ddt=2", "boundary=2", "mdbccorrector=1"], [20], 21),
    ("db2d_ver_wend_art_ddt1_mdbc", "dambreak2d", 3000, ["tstep=1", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "boundary=2", "verletsteps=10"], [25], 26),
    # in-solver gauges of the reference (JGaugeSystem: velocity, swl, maxz, force; computed every step), dumped by the harness at the
    # initial state and at the listed steps; "gauges:" injects the <special><gauges> block below into Case.xml
    ("db_sym_wend_art_ddt2_gauges", "dambreak", 3000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2", "gauges:"], [0, 20, 60], 61),
    ("db_ver_wend_art_ddt2_gauges", "dambreak", 6000, ["tstep=1", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2", "verletsteps=10", "gauges:"], [35], 36),
    # moving boundaries (a plate right of the water column): gate lifted at constant speed for 0.004 s then at rest (mvrect -> wait: the
    # chain ends inside a step), flap oscillating about its lower edge (mvrotsinu: matrix motion) and piston oscillating along x
    # (mvrectsinu); the harness logs what JDsMotion computed at every step (motionlog) next to the usual stages
    ("mv_sym_gate_rect_wait", "movgate", 3000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=2"], [0, 9, 30], 31),
    ("mv_ver_flap_rotsinu", "movflap", 3000, ["tstep=1", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "verletsteps=10"], [0, 25], 26),
    ("mv_sym_piston_rectsinu_cubic", "movpiston", 3000, ["tstep=2", "tkernel=1", "tvisco=1", "visco=0.1", "ddt=2", "shifting=3", "shiftcoef=-2"], [25], 26),
    ("mv2d_sym_flap_rotsinu_mdbc", "movflap2d", 3000, ["tstep=2", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "boundary=2"], [25], 26),
    ("db_ver_wend_art_ddt1_rhopout", "dambreak", 3000, ["tstep=1", "tkernel=2", "tvisco=1", "visco=0.1", "ddt=1", "verletsteps=3", "xml:RhopOutMax=1002.14", "xml:RhopOutMin=999.2"], [12, 30], 31),
]


GAUGES_XML = """    <special>
      <gauges>
        <default>
          <savevtkpart value="false" />
          <computedt value="0" />
          <computetime start="0" end="10" />
          <output value="true" />
          <outputdt value="0" />
          <outputtime start="0" end="10" />
        </default>
        <velocity name="Vcol"><point x="0.2" y="0.33" z="0.1" /></velocity>
        <velocity name="Vtoe"><point x="0.38" y="0.30" z="0.06" /></velocity>
        <velocity name="Vdry"><point x="0.9" y="0.33" z="0.05" /></velocity>
        <swl name="Hcol"><point0 x="0.2" y="0.33" z="0.0" /><point2 x="0.2" y="0.33" z="0.4" /><pointdp coefdp="0.5" /><masslimit coef="0.5" /></swl>
        <swl name="Front"><point0 x="0.05" y="0.33" z="0.03" /><point2 x="1.55" y="0.33" z="0.03" /><pointdp coefdp="0.5" /></swl>
        <maxz name="Zcol"><point0 x="0.3" y="0.33" z="0.0" /><height value="0.4" /><distlimit coefdp="3" /></maxz>
        <maxz name="Zdry"><point0 x="1.2" y="0.33" z="0.0" /><height value="0.4" /><distlimit coefh="2" /></maxz>
        <force name="Ftank"><target mkbound="0" /></force>
      </gauges>
    </special>
"""


def inject_gauges(xmlfile):
    """<special><gauges> of GAUGES_XML right behind <execution> (the synthetic single-phase cases have no <special> of their own)."""
    xml = open(xmlfile).read()
    assert "<special>" not in xml and "<execution>" in xml
    open(xmlfile, "w").write(xml.replace("<execution>\n", "<execution>\n" + GAUGES_XML, 1))


def run(name, kind, nptarget, opts, steps, nsteps, outdir):
    tmp = tempfile.mkdtemp(prefix="golden_")
    try:
        subprocess.check_call([MKCASE, kind, str(nptarget), tmp, "Case"] + [o for o in opts if not o.startswith("xml:") and o != "gauges:"], stdout=subprocess.DEVNULL)
        xmlfile = os.path.join(tmp, "Case.xml")
        if "gauges:" in opts:
            inject_gauges(xmlfile)
        xml = open(xmlfile).read()
        for o in opts:
            if o.startswith("xml:"):
                k, v = o[4:].split("=")
                import re
                xml, n = re.subn(r'(<parameter key="%s" value=")[^"]*(")' % k, r"\g<1>%s\g<2>" % v, xml)
                assert n == 1, "parameter %s not found in Case.xml" % k
        open(xmlfile, "w").write(xml)
        cmd = [HARNESS_NN if kind.startswith("nn") else HARNESS, "Case", "out", "-cpu", "-ompthreads:1", "-sv:none", "-hdir:" + os.path.join(tmp, "dump"),
               "-hsteps:" + ",".join(str(s) for s in steps), "-hnsteps:%d" % nsteps]
        cmd += ["-cellmode:half"] if "cellmode=2" in opts else []
        r = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            print(r.stdout[-3000:])
            raise RuntimeError("reference harness failed for " + name)
        cd = dsph.load_case(tmp, "Case")
        if "cellmode=2" in opts:
            cd.cfg.cellmode = 2
        pack = {"config": np.frombuffer(bytes(cd.cfg), np.uint8).copy(), "case/idp": cd.idp, "case/pos": cd.pos,
                "case/velrhop": cd.velrhop, "case/code": cd.codes(), "case/npb": np.uint32(cd.npb), "steps": np.array(steps, np.uint32),
                "nsteps": np.uint32(nsteps), "case/caseposmin": np.array(cd.caseposmin), "case/caseposmax": np.array(cd.caseposmax)}
        if "boundary=2" in opts:
            pack["case/normals"] = dsph.case_normals(tmp, "Case")
        if "gauges:" in opts or kind.startswith("mov"):
            pack["case/xml"] = np.frombuffer(open(xmlfile, "rb").read(), np.uint8).copy()
        for f in sorted(os.listdir(os.path.join(tmp, "dump"))):
            d = read_dump(os.path.join(tmp, "dump", f))
            stem = f[:-4]
            for k, v in d.items():
                pack[stem + "/" + k] = v
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **pack)
        print("%-32s np=%d  %.1f KB" % (name, cd.np_, os.path.getsize(os.path.join(outdir, name + ".npz")) / 1024))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    if not os.path.exists(HARNESS):
        sys.exit("oracle/_ref/ref_harness52 is missing: run `make -C oracle v52` where /root/reference is mounted")
    out = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for c in CASES:
        if not only or c[0] in only:
            run(*c, outdir=out)
