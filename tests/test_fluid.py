"""Fluid forces of mj_passive (mjOption density / viscosity / wind, MuJoCo's inertia-box model) and body gravcomp on the path of
Physics::step (reference src/physics.rs:44-46): compiler table, closed forms for the oracle, refusals, and - with zoo_r -
the host instantiation of the stage templates and (GPU) every kernel family against the oracle. tests/test_golden.py pins
the oracle on zoo_r against the dense checker, which derives the same forces from body velocities and exact Jacobians."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

BOX = """<mujoco><option timestep="0.002" density="{rho}" viscosity="{mu}" wind="{wind}" gravity="0 0 {g}"/><worldbody>
<body name="b" pos="0 0 1"><freejoint/><geom type="box" size="0.1 0.2 0.3" density="500"/></body></worldbody></mujoco>"""


def test_compiled_box_table():
    """A uniform box IS its own inertia box: sides 0.2 x 0.4 x 0.6, equivalent diameter 0.4."""
    m = ox.Model.from_xml_string(BOX.format(rho=1.2, mu=0.3, wind="0.5 0 0", g=0))
    assert m.nfluid == m.nbody == 2 and m.density == 1.2 and m.viscosity == 0.3 and list(m.wind) == [0.5, 0, 0]
    f = m.body_fluid.reshape(2, 11)
    assert not f[0].any()
    a, b, c, d = 0.2, 0.4, 0.6, 0.4
    want = [np.pi * d ** 3 * 0.3, 3 * np.pi * d * 0.3, 0.6 * b * c, 0.6 * a * c, 0.6 * a * b,
            1.2 * a * (b ** 4 + c ** 4) / 64, 1.2 * b * (a ** 4 + c ** 4) / 64, 1.2 * c * (a ** 4 + b ** 4) / 64, 0.5, 0, 0]
    assert np.allclose(f[1], want, rtol=1e-12)
    assert ox.Model.from_xml_string(BOX.format(rho=0, mu=0, wind="1 1 1", g=0)).nfluid == 0      # no medium, no table


def test_drag_on_a_free_box_closed_form():
    """Body axes = world axes: force_k = -3 pi d mu v_k - rho/2 A_k |v_k| v_k on the velocity relative to the wind, torque likewise."""
    rho, mu, wind = 1.2, 0.3, np.array([0.5, -0.2, 0.1])
    m = ox.Model.from_xml_string(BOX.format(rho=rho, mu=mu, wind="0.5 -0.2 0.1", g=0))
    v, w = np.array([1.0, -2.0, 0.5]), np.array([0.3, -0.4, 2.0])
    od = OracleData(m)
    od.field("qvel")[:] = np.concatenate([v, w])
    od.forward()
    s, d = np.array([0.2, 0.4, 0.6]), 0.4
    rel = v - wind
    area = np.array([s[1] * s[2], s[0] * s[2], s[0] * s[1]])
    quart = np.array([s[1] ** 4 + s[2] ** 4, s[0] ** 4 + s[2] ** 4, s[0] ** 4 + s[1] ** 4])
    force = -3 * np.pi * d * mu * rel - 0.5 * rho * area * np.abs(rel) * rel
    torque = -np.pi * d ** 3 * mu * w - rho * s * quart * np.abs(w) * w / 64
    assert np.allclose(od.field("qfrc_passive"), np.concatenate([force, torque]), rtol=1e-12, atol=1e-15)
    # rotated body: the drag is computed in the body's inertial frame - rotating body, velocity and wind together rotates the force
    q = np.array([0.8, 0.2, -0.4, 0.4]); q /= np.linalg.norm(q)
    R = np.array([[1 - 2 * (q[2] ** 2 + q[3] ** 2), 2 * (q[1] * q[2] - q[0] * q[3]), 2 * (q[1] * q[3] + q[0] * q[2])],
                  [2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[1] ** 2 + q[3] ** 2), 2 * (q[2] * q[3] - q[0] * q[1])],
                  [2 * (q[1] * q[3] - q[0] * q[2]), 2 * (q[2] * q[3] + q[0] * q[1]), 1 - 2 * (q[1] ** 2 + q[2] ** 2)]])
    wr = R @ wind
    m2 = ox.Model.from_xml_string(BOX.format(rho=rho, mu=mu, wind=" ".join(repr(float(x)) for x in wr), g=0))
    od2 = OracleData(m2)
    od2.field("qpos")[3:7] = q
    od2.field("qvel")[:] = np.concatenate([R @ v, w])        # free joint: linear velocity in world axes, angular in body axes
    od2.forward()
    assert np.allclose(od2.field("qfrc_passive"), np.concatenate([R @ force, torque]), rtol=1e-11, atol=1e-14)


def test_terminal_velocity():
    """A ball sinking in a viscous, dense medium settles where weight = Stokes drag + quadratic drag (the medium carries no buoyancy)."""
    xml = """<mujoco><option timestep="0.002" density="50" viscosity="2"/><worldbody>
    <body pos="0 0 5"><freejoint/><geom type="sphere" size="0.05" density="2000" contype="0" conaffinity="0"/></body></worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m)
    for _ in range(12000):
        od.step()
    vz = -od.field("qvel")[2]
    mass = float(m.body_mass[1])
    side = np.sqrt(6 * 2 * (0.4 * mass * 0.05 ** 2 - 0.5 * 0.4 * mass * 0.05 ** 2) / mass)      # sqrt(6 (I + I - I) / m), I = 2/5 m r^2
    assert abs(side - np.sqrt(2.4) * 0.05) < 1e-12
    balance = 3 * np.pi * side * 2 * vz + 0.5 * 50 * side ** 2 * vz ** 2 - mass * 9.81
    assert vz > 0.5 and abs(balance) < 1e-9 * mass * 9.81
    assert abs(od.field("qacc")[2]) < 1e-9


def test_flag_and_refusals():
    m = ox.Model.from_xml_string(BOX.format(rho=5, mu=1, wind="0 0 0", g=0).replace("/><worldbody>", '><flag passive="disable"/></option><worldbody>'))
    od = OracleData(m)
    od.field("qvel")[:] = [1, 2, 3, 1, 2, 3]
    od.forward()
    assert not od.field("qfrc_passive").any()
    with pytest.raises(ox.Error, match="implicitfast"):
        ox.Model.from_xml_string(BOX.format(rho=5, mu=0, wind="0 0 0", g=0).replace("<option ", '<option integrator="implicitfast" '))
    with pytest.raises(ox.Error, match="fluidshape"):
        ox.Model.from_xml_string(BOX.format(rho=5, mu=0, wind="0 0 0", g=0).replace('density="500"', 'density="500" fluidshape="ellipsoid"'))
    with pytest.raises(ox.Error, match="density"):
        ox.Model.from_xml_string(BOX.format(rho=-1, mu=0, wind="0 0 0", g=0))


def test_gravcomp_closed_forms():
    """body gravcomp = c adds -c * m * gravity at the body's com (mj_passive): c = 1 hovers, 0.5 halves the fall, 2 rises; a chain
    whose bodies are all fully compensated is in equilibrium in every pose; the gravity and passive flags switch it off."""
    free = '<mujoco><option {o}/><worldbody><body pos="0 0 1" gravcomp="{c}"><freejoint/><geom type="box" size="0.1 0.2 0.3"/></body></worldbody></mujoco>'
    for c, az in ((1, 0.0), (0.5, -4.905), (2, 9.81), (0, -9.81)):
        m = ox.Model.from_xml_string(free.format(o='gravity="0 0 -9.81"', c=c))
        assert m.ngravcomp == (m.nbody if c else 0)
        od = OracleData(m); od.forward()
        assert np.allclose(od.field("qacc"), [0, 0, az, 0, 0, 0], atol=1e-12)
    m = ox.Model.from_xml_string(free.format(o='gravity="1 2 -9.81"', c=1).replace("/><worldbody>", '><flag passive="disable"/></option><worldbody>'))
    od = OracleData(m); od.forward()
    assert np.allclose(od.field("qacc")[:3], [1, 2, -9.81], atol=1e-12)
    chain = """<mujoco><worldbody><body pos="0 0 1" gravcomp="1"><joint axis="0 1 0"/><geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.03"/>
    <body pos="0.4 0 0" gravcomp="1"><joint type="ball"/><geom type="capsule" fromto="0 0 0 0.2 0.1 0.1" size="0.02"/>
    <body pos="0.2 0.1 0.1" gravcomp="{c}"><joint type="slide" axis="1 1 0"/><geom size="0.05"/></body></body></body></worldbody></mujoco>"""
    m = ox.Model.from_xml_string(chain.format(c=1))
    qpos, _ = random_state(m, 3, seed=5)
    for e in range(3):
        od = OracleData(m); od.field("qpos")[:] = qpos[e]; od.forward()
        assert np.abs(od.field("qacc")).max() < 1e-10 and np.abs(od.field("qfrc_bias")).max() > 0.1
        assert np.allclose(od.field("qfrc_passive"), od.field("qfrc_bias"), atol=1e-12)
    m = ox.Model.from_xml_string(chain.format(c=0))                                 # the last body uncompensated: it alone pulls
    od = OracleData(m); od.field("qpos")[:] = qpos[0]; od.forward()
    assert np.abs(od.field("qacc")).max() > 0.1


def test_swimmer_is_propelled_by_the_medium():
    """zoo_r's swimmer hangs on planar root joints: without a medium an internal torque cannot move its centre of mass, with
    one it does."""
    base = ZOO["zoo_r"]
    for j in ("rootx", "rooty", "rootz"):                            # frictionless, inertia-free root joints
        base = base.replace(f'<joint name="{j}"', f'<joint name="{j}" damping="0" armature="0"')
    base = base.replace('<option timestep="0.003"', '<option integrator="RK4" timestep="0.003"')    # Euler drifts O(h) in momentum
    dry = base.replace('density="900" viscosity="0.2" wind="0.3 -0.1 0.05"', "")
    still = base.replace('wind="0.3 -0.1 0.05"', "")
    moved = []
    for xml in (dry, still):
        m = ox.Model.from_xml_string(xml)
        od = OracleData(m)
        for s in range(400):
            od.field("ctrl")[:] = [np.sin(0.05 * s), np.sin(0.05 * s - 1.5)]
            od.step()
        sm = od.field("subtree_com").reshape(-1, 3)[1]
        moved.append(sm.copy())
    m = ox.Model.from_xml_string(dry)
    od = OracleData(m); od.forward()
    start = od.field("subtree_com").reshape(-1, 3)[1].copy()
    assert np.linalg.norm(moved[0][:2] - start[:2]) < 1e-4          # momentum conservation in the plane (1.1e-5 of RK4 truncation; 1.2e-2 with Euler)
    assert np.linalg.norm(moved[1][:2] - start[:2]) > 1e-3


def test_fluid_host_instantiation_long_horizon():
    m = ox.Model.from_xml_string(ZOO["zoo_r"])
    nenv, nsteps = 4, 300
    qpos, qvel = random_state(m, nenv, seed=211)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel * 3)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e] * 3
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        for f in ("qpos", "qvel", "qacc", "qfrc_passive", "sensordata"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-7, (f, e)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,specialize,precision,tol", [("fused", 0, "f64", 1e-7), ("staged", 0, "f64", 1e-7), ("fused", 2, "f64", 1e-7),
                                                           ("coop", 0, "f64", 1e-7), ("fused", 2, "f32", 2e-2)])
def test_fluid_gpu_vs_oracle(mode, specialize, precision, tol):
    m = ox.Model.from_xml_string(ZOO["zoo_r"])
    nenv, nsteps = 64, 120
    qpos, qvel = random_state(m, nenv, seed=223)
    b = ox.BatchedPhysics(m, nenv, precision=precision, mode=mode, specialize=specialize)
    if specialize == 2:
        assert "jit" in b.kernel_name()
    b.set("qpos", qpos); b.set("qvel", qvel * 3); b.ctrl_philox(True, SEED)
    b.step(1); b.sync()
    ref1 = []
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e] * 3
        od.fill_ctrl_philox(e, 0); od.step()
        ref1.append(od.field("qacc").copy()); ods.append(od)
    assert rel_err(b.get("qacc"), np.stack(ref1)) <= (1e-9 if precision == "f64" else 2e-3)
    if mode != "fused" or specialize == 0:     # the specialised kernels keep derived fields in registers (include/ox_b200.h)
        ref = []
        for e in range(nenv):
            od = OracleData(m)
            od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e] * 3
            od.fill_ctrl_philox(e, 0); od.forward()
            ref.append(od.field("qfrc_passive").copy())
        assert np.abs(np.stack(ref)).max() > 1.0                                  # the medium really acts
    b.step(nsteps - 1); b.sync()
    for e, od in enumerate(ods):
        for s in range(1, nsteps):
            od.fill_ctrl_philox(e, s); od.step()
    for f in ("qpos", "sensordata"):
        assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= tol, f
    assert int(b.diverged().sum()) == 0
