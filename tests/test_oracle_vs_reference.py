"""CPU checker (oracle/sph_oracle.cpp) against dumps of the reference's own JSphCpuSingle (tests/golden/*.npz).

This is what pins the oracle: the reference ships no golden vectors, so the fixtures are outputs of the reference
solver itself (generator: tests/golden/make_golden.py).  Runs without a GPU."""
import pytest

import parity
from parity import Golden

# the reference is built with -ffast-math, the checker is not: ~2e-6 of relative L2 is the reference's own noise floor
TOL = dict(default=2e-5, ace=1e-5, ar=1e-5, dt=1e-6, viscdt=1e-5, acemax=1e-5, velmax=1e-6, pos=1e-8, state=1e-6)
# periodic box: uniform density, Mach 0.01 -> |ace| ~ 1e-2 m/s^2 is the residue of pressure terms that cancel to ~1e-5
# of their magnitude, so the fast-math/strict-math difference of the two CPU builds already reaches ~1e-5 relative L2.
TOL_PB = dict(TOL, ace=1e-4)


def tol_for(name):
    return TOL_PB if name.startswith("pb_") else TOL


@pytest.mark.parametrize("name", parity.GOLDEN_CASES)
def test_oracle_stages(oracle_lib, name):
    g = Golden(name)
    for step in g.steps:
        res = parity.run_step_stages(oracle_lib, g, step, tol_for(name))
        print(name, step, res)


@pytest.mark.parametrize("name", parity.GOLDEN_CASES)
def test_oracle_trajectory(oracle_lib, name):
    g = Golden(name)
    out = parity.run_trajectory(oracle_lib, g)
    print(name, out)
    assert out["np"][0] == out["np"][1]
    assert out["time"] < 1e-6
    assert out["pos"] < 1e-6 and out["velrhop"] < 1e-4
