"""Flow measurements used by the 1-s dam-break validation (north_star: gauge pressures and surge-front position within 1 %).

A small stand-in for the reference's MeasureTool (`-points ... -vars:+press,+kcorr`,
R/examples/main/01_DamBreak/xCaseDambreak_linux64_CPU.sh:76-85): pressure at a probe point is the Shepard-corrected
Wendland interpolation over the fluid particles, P(x) = sum_j P_j W(x-x_j) m/rho_j / sum_j W(x-x_j) m/rho_j with
P_j = B((rho_j/rho0)^gamma - 1).  The same code measures the reference's PART files and ours."""
import numpy as np


def wendland(q, h):
    aw = 0.41778 / h ** 3                     # 3-D Wendland constant, S/FunSphKernel.h:199-210
    w = aw * (1.0 - 0.5 * q) ** 4 * (2.0 * q + 1.0)
    return np.where(q < 2.0, w, 0.0)


def gauge_pressure(points, pos, rhop, h, cteb, rhop0, gamma, mass):
    """Shepard-corrected pressure at `points` [k,3] from fluid particles pos [n,3], rhop [n]."""
    press = cteb * ((rhop.astype(np.float64) / rhop0) ** gamma - 1.0)
    out = np.zeros(len(points)); ksum = np.zeros(len(points))
    for k, pt in enumerate(np.asarray(points, np.float64)):
        d = pos - pt
        near = np.all(np.abs(d) <= 2.0 * h, axis=1)
        if not near.any():
            continue
        r = np.sqrt((d[near] ** 2).sum(axis=1))
        w = wendland(r / h, h) * (mass / rhop[near])
        ksum[k] = w.sum()
        out[k] = (w * press[near]).sum() / ksum[k] if ksum[k] > 0 else 0.0
    return out, ksum


def gauge_velocity(points, pos, vel, rhop, h, mass):
    """Shepard-corrected velocity at `points` [k,3] from fluid particles (JGaugeVelocity interpolation / kernel sum)."""
    out = np.zeros((len(points), 3))
    for k, pt in enumerate(np.asarray(points, np.float64)):
        d = pos - pt
        near = np.all(np.abs(d) <= 2.0 * h, axis=1)
        if not near.any():
            continue
        r = np.sqrt((d[near] ** 2).sum(axis=1))
        w = wendland(r / h, h) * (mass / rhop[near])
        if w.sum() > 0:
            out[k] = (w[:, None] * vel[near]).sum(axis=0) / w.sum()
    return out


def surge_front(pos_fluid, zmax):
    """x of the advancing front along the bed: the 99.5th percentile of x of the fluid particles below zmax."""
    sel = pos_fluid[:, 2] <= zmax
    return float(np.percentile(pos_fluid[sel, 0], 99.5)) if sel.any() else float("nan")


def column_height(pos_fluid, xmax):
    """free-surface height at the upstream wall: 99th percentile of z of the fluid particles with x<=xmax."""
    sel = pos_fluid[:, 0] <= xmax
    return float(np.percentile(pos_fluid[sel, 2], 99.0)) if sel.any() else float("nan")


def measure_state(pos, rhop, idp, cfg, gauges):
    """Front, column height and gauge pressures of one saved state (all particles; fluid = id >= casenfixed)."""
    fl = idp >= cfg.casenfixed
    pf = pos[fl]; rf = rhop[fl]
    h = float(cfg.kernelh)
    p, ks = gauge_pressure(gauges, pf, rf, h, float(cfg.cteb), float(cfg.rhopzero), float(cfg.gamma), float(cfg.massfluid))
    return dict(front=surge_front(pf, 3.0 * cfg.dp), height=column_height(pf, 0.05), press=p, ksum=ks)
