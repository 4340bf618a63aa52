"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol declared in include/*.h, the case
tooling round-trips, the bi4 files are accepted by the reference solver (when it was built here), and compute entry
points fail loudly without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import dualsphysics_multilayer_b200 as dsph
from dualsphysics_multilayer_b200 import capi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dsph_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dsph_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(dsph.LIBPATH)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_struct_sizes_match_header():
    # sizeof checked through a tiny C program compiled against the header
    src = r'''
    #include "dsph_b200.h"
    #include <stdio.h>
    int main(){ printf("%zu %zu %zu %zu %zu %zu\n",sizeof(dsph_config),sizeof(dsph_divinfo),sizeof(dsph_forceres),sizeof(dsph_timers),sizeof(dsph_case),sizeof(dsph_phase)); return 0; }
    '''
    import tempfile
    d = tempfile.mkdtemp()
    open(os.path.join(d, "s.c"), "w").write(src)
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
    out = subprocess.check_output([os.path.join(d, "s")]).split()
    got = [int(x) for x in out]
    want = [C.sizeof(capi.Config), C.sizeof(capi.DivInfo), C.sizeof(capi.ForceRes), C.sizeof(capi.Timers), C.sizeof(capi.Case), C.sizeof(capi.Phase)]
    assert got == want


@pytest.mark.parametrize("kind,n", [("dambreak", 5000), ("periodicbox", 1000), ("nn2layer", 4000)])
def test_case_make_save_load_roundtrip(tmp_path, kind, n):
    cd = dsph.make_case(kind, n)
    assert cd.np_ > 0 and cd.npb == cd.cfg.casenfixed
    assert np.array_equal(cd.idp, np.arange(cd.np_, dtype=np.uint32))
    dsph.save_case(cd, str(tmp_path), "Case")
    cd2 = dsph.load_case(str(tmp_path), "Case")
    assert cd2.np_ == cd.np_ and cd2.npb == cd.npb
    assert np.array_equal(cd2.pos, cd.pos) and np.array_equal(cd2.velrhop, cd.velrhop) and np.array_equal(cd2.idp, cd.idp)
    for f in ("kernelh", "cteb", "gamma", "rhopzero", "massfluid", "massbound", "dp", "cflnumber", "tstep", "tkernel", "tvisco",
              "tdensity", "shiftmode", "perix", "periy", "periz", "nn_on", "nn_nphases"):
        assert getattr(cd2.cfg, f) == getattr(cd.cfg, f), f
    assert np.allclose(list(cd2.cfg.maprealposmin), list(cd.cfg.maprealposmin), rtol=0, atol=1e-12)
    assert np.allclose(list(cd2.cfg.maprealposmax), list(cd.cfg.maprealposmax), rtol=0, atol=1e-12)
    assert cd2.fluidblocks == cd.fluidblocks


def test_compute_fails_loudly_without_gpu():
    if dsph.device_count() > 0:
        pytest.skip("a CUDA device is present")
    cd = dsph.make_case("dambreak", 3000)
    with pytest.raises(dsph.SphError) as e:
        dsph.create_context(cd)
    assert "no CPU fallback" in str(e.value)


def test_reference_solver_accepts_our_case(tmp_path):
    ref = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.2CPU_linux64")
    if not os.path.exists(ref):
        pytest.skip("reference CPU solver not built here")
    cd = dsph.make_case("dambreak", 3000)
    dsph.save_case(cd, str(tmp_path), "Case")
    r = subprocess.run([ref, "Case", "out", "-cpu", "-ompthreads:2", "-sv:none", "-nsteps:3"], cwd=str(tmp_path),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert "Finished execution (code=0)" in r.stdout, r.stdout[-2000:]
    assert "Particles of simulation (initial): %d" % cd.np_ in r.stdout


def test_plain_c_host_program_uses_the_host_side_of_the_abi(tmp_path):
    """A C99 program (no C++, no Python) linked against libdsph_b200.so: the slab cuts of dsph_mgpu_balanced_cuts and the motion laws of a
    case through dsph_case_motion_open / process / close -- the calls a C host makes around the device entry points.  No GPU needed."""
    mkcase = os.path.join(ROOT, "dualsphysics_multilayer_b200", "bin", "mkcase")
    subprocess.check_call([mkcase, "movpiston", "3000", str(tmp_path), "Case"], stdout=subprocess.DEVNULL)
    src = r'''
    #include "dsph_b200.h"
    #include <stdio.h>
    #include <math.h>
    int main(int argc,char **argv){
      double work[40]; uint32_t cuts[5]; int i;
      for(i=0;i<40;i++)work[i]=(i<10? 100.0: 5.0);
      if(dsph_mgpu_balanced_cuts(work,40,4,3,cuts)!=DSPH_OK){ printf("cuts failed: %s\n",dsph_last_error()); return 1; }
      if(cuts[0]!=0 || cuts[4]!=40)return 2;
      if(dsph_mgpu_balanced_cuts(work,5,4,3,cuts)==DSPH_OK)return 3;            /* too many ranks for 5 columns */
      dsph_motionset *ms=0; uint32_t nobj=0;
      if(dsph_case_motion_open(argv[1],"Case",&ms,&nobj)!=DSPH_OK){ printf("motion failed: %s\n",dsph_last_error()); return 4; }
      if(nobj!=1)return 5;
      dsph_motiondata d; int active=0; double t=0,dt=1e-3,x=0;
      for(i=0;i<50;i++){ if(dsph_case_motion_process(ms,t,dt,&d,&active)!=DSPH_OK)return 6; if(!active || d.type!=DSPH_MOTION_LINEAR)return 7; x+=d.linmov[0]; t+=dt; }
      /* mvrectsinu of mkcase: x(t) = 0.1 sin(2 pi 2 t) */
      if(fabs(x-0.1*sin(2*3.14159265358979323846*2*t))>1e-12){ printf("x=%.15g\n",x); return 8; }
      dsph_case_motion_close(ms);
      printf("c host ok %u %u %u\n",cuts[1],cuts[2],cuts[3]);
      return 0;
    }
    '''
    open(str(tmp_path / "h.c"), "w").write(src)
    libdir = os.path.dirname(dsph.LIBPATH)
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(tmp_path / "h"), str(tmp_path / "h.c"),
                           "-L", libdir, "-ldsph_b200", "-lm", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(tmp_path / "h"), str(tmp_path)], stdout=subprocess.PIPE, text=True)
    assert out.returncode == 0 and "c host ok" in out.stdout, (out.returncode, out.stdout)
