"""N3 row of SURVEY 8f: elliptic friction cones (`<option cone="elliptic"/>`) in the primal solvers (Newton, CG).

The per-contact cost (three zones in scaled coordinates, oracle `ellipticCost` / kernel `elliptic_cost`) is pinned as the exact
convex conjugate of the regularised dual problem: s(x) = max over the TRUE friction cone of (-1/2 f'Rf - f'x), maximised
numerically by projected gradient on the second-order cone. Coulomb's law pins the behaviour (a block slides down a slope with g (sin t - mu cos t) and sticks when
tan t < mu). tests/test_golden.py checks zoo_m / zoo_n against the dense checker, which differentiates the objective with autograd."""
import ctypes as C

import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

SLOPE = """<mujoco><compiler angle="radian"/><option timestep="0.002" cone="{cone}" impratio="{imp}" tolerance="1e-12"/><worldbody>
<geom type="plane" size="5 5 0.1" euler="0 {t} 0" friction="{mu} 0.005 0.0001"/>
<body pos="{px} 0 {pz}" euler="0 {t} 0"><freejoint/><geom type="box" size="0.1 0.1 0.05" friction="{mu} 0.005 0.0001"/></body></worldbody></mujoco>"""


def test_cost_is_the_conjugate_of_the_regularised_cone_problem():
    """One contact at rest on a plane, read its rows back and compare efc_force with the maximiser of -1/2 f'Rf - f'x over
    {f_n >= 0, |f_t / friction| <= f_n}, found by projected gradient for the SAME x = J qacc - aref."""
    xml = """<mujoco><option cone="elliptic" impratio="2"/><worldbody><geom type="plane" size="1 1 0.1"/>
    <body pos="0 0 0.098"><freejoint/><geom type="sphere" size="0.1" condim="6" friction="0.7 0.05 0.02"/></body></worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    rng = np.random.default_rng(5)
    zones = set()
    for trial in range(40):
        od = OracleData(m)
        od.field("qvel")[:] = rng.normal(0, [0.5, 0.5, 0.3, 3, 3, 3]) * (1e-3 if trial % 2 else 1.0)   # near rest: inside the cone (stick)
        od.field("qpos")[2] += rng.uniform(-0.004, 0.002)
        od.forward()
        if od.int("nefc") != 6:
            continue
        J = od.field("efc_J")[:36].reshape(6, 6)
        D, aref, f = od.field("efc_D")[:6], od.field("efc_aref")[:6], od.field("efc_force")[:6]
        x = J @ od.field("qacc") - aref
        R = 1 / D
        fri = np.asarray(m.pair_friction[:5])          # max of the two geoms: (1, 1, 0.05, 0.02, 0.02) - the plane's default sliding friction is 1
        obj = lambda g: 0.5 * g @ (R * g) + g @ x
        # minimise obj over the cone by projected gradient in the coordinates g_j = f_j / friction_j, where the cone is the standard
        # second-order cone |g_t| <= g_0 and the projection has a closed form
        sc = np.r_[1.0, fri]
        Rg, xg = R * sc * sc, x * sc
        g = np.zeros(6)
        step = 1 / Rg.max()
        for _ in range(20000):
            y = g - step * (Rg * g + xg)
            nt = np.linalg.norm(y[1:])
            if nt <= y[0]:
                g = y
            elif nt <= -y[0]:
                g = np.zeros(6)
            else:
                a = 0.5 * (y[0] + nt)
                g = np.r_[a, a * y[1:] / nt]
        class best: fun = obj(g * sc)
        assert abs(obj(f) - best.fun) <= 1e-6 * max(1.0, abs(best.fun)), (trial, obj(f), best.fun)
        assert np.sum((f[1:] / fri) ** 2) <= f[0] ** 2 * (1 + 1e-9) and f[0] >= 0          # the force the solver reports is IN the cone
        zones.add("rest" if np.allclose(f, 0) else ("stick" if np.sum((f[1:] / fri) ** 2) < 0.999 * f[0] ** 2 else "slide"))
    assert {"stick", "slide"} <= zones, zones


@pytest.mark.parametrize("solver", ["Newton", "CG"])
def test_coulomb_sliding_and_sticking(solver):
    g, t = 9.81, 0.5
    for mu, slides in ((0.3, True), (0.8, False)):
        xml = SLOPE.format(cone="elliptic", imp=1, t=t, mu=mu, px=0.0497 * np.sin(t), pz=0.0497 * np.cos(t)).replace('cone="elliptic"', f'cone="elliptic" solver="{solver}"')
        od = OracleData(ox.Model.from_xml_string(xml))
        for _ in range(150):
            od.step()
        v0 = od.field("qvel")[:3].copy()
        for _ in range(250):
            od.step()
        acc = np.linalg.norm(od.field("qvel")[:3] - v0) / (250 * 0.002)
        if slides:   # (the sliding block chatters on its soft contacts, so the instantaneous contact count is not asserted)
            assert abs(acc - g * (np.sin(t) - mu * np.cos(t))) < 0.02 * g, (mu, acc)      # Coulomb friction at exactly mu N
        else:
            assert od.int("ncon") == 4 and od.int("nefc") == 12
            assert acc < 0.01 * g and np.linalg.norm(od.field("qvel")[:3]) < 5e-3          # inside the cone: it stays put


def test_compiler_and_refusals():
    m = ox.Model.from_xml_string(ZOO["zoo_m"])
    assert m.cone == 1
    for extra in ('solver="PGS"', 'noslip_iterations="2"'):
        with pytest.raises(ox.MjsError, match="elliptic"):
            ox.Model.from_xml_string(f'<mujoco><option cone="elliptic" {extra}/><worldbody><body><freejoint/><geom size="0.1"/></body></worldbody></mujoco>')


@pytest.mark.parametrize("name", ["zoo_m", "zoo_n"])
def test_host_instantiation_matches_oracle(name):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 5, 200
    qpos, qvel = random_state(m, nenv, seed=79)
    if name == "zoo_m":
        qvel[:, 3:6] *= 40; qvel[:, 9:12] *= 40
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    rows = 0
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        rows += od.int("nefc")
        for f in ("qpos", "qvel", "sensordata"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-6, (f, e)
    assert rows > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zoo_m", "zoo_n"])
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("fused", 2)])
def test_gpu_vs_oracle(name, mode, specialize):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 64, 120
    qpos, qvel = random_state(m, nenv, seed=83)
    if name == "zoo_m":
        qvel[:, 3:6] *= 40; qvel[:, 9:12] *= 40
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        ods.append(od)
    done = 0
    for upto, tol in ((1, 1e-8), (nsteps, 1e-5)):
        b.step(upto - done); b.sync()
        for e, od in enumerate(ods):
            for s in range(done, upto):
                od.fill_ctrl_philox(e, s); od.step()
        done = upto
        for f in ("qpos", "qvel"):
            assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= tol, (f, upto)
    assert sum(od.int("nefc") for od in ods) > 0 and int(b.diverged().sum()) == 0
    with pytest.raises(ox.Error, match="coop"):
        ox.BatchedPhysics(m, 32, precision="f64", mode="coop")
