"""Live runs of the REFERENCE CPU solver on the GPU box: oracle/_ref/ref_harness52 -- the reference's own
JSphCpuSingle compiled by oracle/Makefile, subclassed by oracle/ref_harness.cpp to dump intermediate state -- travels with the
repository snapshot, so the -m gpu tests can pin the CUDA path to the reference at sizes far beyond the committed 3 k-particle
fixtures: the C1 configuration (~170 k particles, 24 % boundary) and DB(2 M).  Nothing here reads /root/reference."""
import os
import shutil
import subprocess
import tempfile

import numpy as np

import dualsphysics_multilayer_b200 as dsph
from dumpio import read_dump
from parity import Golden

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness52")
MKCASE = os.path.join(ROOT, "dualsphysics_multilayer_b200", "bin", "mkcase")


def available():
    return os.path.exists(HARNESS) and os.path.exists(MKCASE)


def live_golden(name, kind, np_target, opts, steps, nsteps, threads=None):
    """mkcase -> reference harness (OpenMP: per-particle sums are sequential in JSphCpu, so the result does not depend on the
    number of threads) -> the dumped stages of `steps` as a Golden."""
    threads = threads or min(os.cpu_count() or 1, 32)
    tmp = tempfile.mkdtemp(prefix="live_ref_")
    try:
        subprocess.check_call([MKCASE, kind, str(np_target), tmp, "Case"] + list(opts), stdout=subprocess.DEVNULL)
        cmd = [HARNESS, "Case", "out", "-cpu", "-ompthreads:%d" % threads, "-sv:none", "-hdir:" + os.path.join(tmp, "dump"),
               "-hsteps:" + ",".join(str(s) for s in steps), "-hnsteps:%d" % nsteps]
        r = subprocess.run(cmd, cwd=tmp, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference harness failed: " + r.stdout[-2000:])
        cd = dsph.load_case(tmp, "Case")
        pack = {"config": np.frombuffer(bytes(cd.cfg), np.uint8).copy(), "case/idp": cd.idp, "case/pos": cd.pos, "case/velrhop": cd.velrhop,
                "case/code": cd.codes(), "case/npb": np.uint32(cd.npb), "steps": np.array(steps, np.uint32), "nsteps": np.uint32(nsteps)}
        pack["case/xml"] = np.frombuffer(open(os.path.join(tmp, "Case.xml"), "rb").read(), np.uint8).copy()
        for f in sorted(os.listdir(os.path.join(tmp, "dump"))):
            d = read_dump(os.path.join(tmp, "dump", f))
            for k, v in d.items():
                pack[f[:-4] + "/" + k] = v
        return Golden(name, pack), cd
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
