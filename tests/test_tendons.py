"""N3 row of SURVEY 8f: fixed tendons (length = sum of coef * joint coordinate) with limits, dead-band springs and dampers, and
the tendonpos / tendonvel sensors. Closed forms pin the oracle; tests/test_golden.py pins it on zoo_h against the dense checker;
the host instantiation of the stage templates and (GPU) every kernel family must match the oracle."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

TWO = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 0"/><worldbody>
<body name="a" pos="0 0 1"><joint name="s1" type="slide" axis="1 0 0"/><geom type="sphere" size="0.1" mass="2" contype="0" conaffinity="0"/></body>
<body name="b" pos="0 1 1"><joint name="s2" type="slide" axis="1 0 0"/><geom type="sphere" size="0.1" mass="3" contype="0" conaffinity="0"/></body>
</worldbody><tendon><fixed name="t" {attrs}><joint joint="s1" coef="1"/><joint joint="s2" coef="-2"/></fixed></tendon>
<sensor><tendonpos tendon="t"/><tendonvel tendon="t"/></sensor></mujoco>"""


def test_spring_damper_closed_form_and_dead_band():
    """Two sliding masses, L = q1 - 2 q2: f = -k (L - hi) - b Ldot above the dead band, qfrc = J' f, qacc = qfrc / m."""
    m = ox.Model.from_xml_string(TWO.format(attrs='stiffness="40" damping="3" springlength="-0.1 0.2"'))
    assert (m.ntendon, m.nwrap) == (1, 2) and m.object_id(ox.obj.Tendon, "t").index == 0
    assert abs(m.tendon_invweight0[0] - (1 / 2 + 4 / 3)) < 1e-12            # J M^-1 J'
    for q1, q2, v1, v2 in ((0.5, 0.05, 0.3, -0.2), (-0.4, 0.1, 0.0, 0.5), (0.1, 0.02, 1.0, 0.0)):
        od = OracleData(m)
        od.field("qpos")[:] = [q1, q2]; od.field("qvel")[:] = [v1, v2]
        od.forward()
        L, Ld = q1 - 2 * q2, v1 - 2 * v2
        f = (40 * (0.2 - L) if L > 0.2 else 40 * (-0.1 - L) if L < -0.1 else 0.0) - 3 * Ld
        assert np.allclose(od.field("sensordata"), [L, Ld], atol=1e-15)
        assert np.allclose(od.field("qfrc_passive"), [f, -2 * f], atol=1e-13)
        assert np.allclose(od.field("qacc"), [f / 2, -2 * f / 3], atol=1e-12)
    # springlength unspecified = the length at qpos0, no dead band
    m0 = ox.Model.from_xml_string(TWO.format(attrs='stiffness="40"'))
    assert list(m0.tendon_lengthspring) == [0.0, 0.0]


def test_limit_row_closed_form():
    """A single active tendon-limit row: qacc = a0 + M^-1 J' f with f = D (aref - J a0) / (1 + D A), A = J M^-1 J' = invweight0."""
    m = ox.Model.from_xml_string(TWO.format(attrs='limited="true" range="-0.3 0.4"'))
    od = OracleData(m)
    od.field("qpos")[:] = [0.5, 0.02]          # L = 0.46 > 0.4: upper limit violated by 0.06
    od.field("qvel")[:] = [0.2, 0.0]
    od.forward()
    assert od.int("nefc") == 1 and abs(od.field("efc_pos")[0] - (0.4 - 0.46)) < 1e-15
    J = np.array([-1.0, 2.0])                   # -side * coef, side = +1
    assert np.allclose(od.field("efc_J")[:2], J)
    A, D, aref = float(m.tendon_invweight0[0]), od.field("efc_D")[0], od.field("efc_aref")[0]
    f = D * aref / (1 + D * A)                  # a0 = 0 (no gravity, no passive force)
    assert f > 0 and abs(od.field("efc_force")[0] - f) <= 1e-9 * f
    assert np.allclose(od.field("qacc"), J * f / np.array([2.0, 3.0]), rtol=1e-9)
    od.field("qpos")[:] = [0.1, 0.0]
    od.forward()
    assert od.int("nefc") == 0


def test_tendon_transmission_closed_form():
    """A motor on the tendon L = q1 - 2 q2 with gear 3: generalised force = gear * u * (1, -2); a position servo on it reads the
    tendon length through the gear (actuatorpos / actuatorvel sensors)."""
    xml = TWO.format(attrs="").replace("</tendon>", "</tendon><actuator><motor name='m' tendon='t' gear='3'/>"
                                      "<position name='p' tendon='t' kp='5' kv='0.5' gear='2'/></actuator>")
    xml = xml.replace("<sensor>", "<sensor><actuatorpos actuator='p'/><actuatorvel actuator='p'/>")
    m = ox.Model.from_xml_string(xml)
    assert list(m.actuator_trntype) == [3, 3] and list(m.actuator_trnid) == [0, 0]
    od = OracleData(m)
    q1, q2, v1, v2, u = 0.3, -0.1, 0.2, 0.4, 0.6
    od.field("qpos")[:] = [q1, q2]; od.field("qvel")[:] = [v1, v2]; od.field("ctrl")[:] = [u, 0.25]
    od.forward()
    L, Ld = q1 - 2 * q2, v1 - 2 * v2
    fp = 5 * 0.25 - 5 * (2 * L) - 0.5 * (2 * Ld)              # position servo: kp (ctrl - length) - kv velocity, length = gear * L
    assert np.allclose(od.field("actuator_force"), [u, fp], atol=1e-14)
    gen = 3 * u + 2 * fp
    assert np.allclose(od.field("qfrc_actuator"), [gen, -2 * gen], atol=1e-13)
    assert np.allclose(od.field("sensordata")[:2], [2 * L, 2 * Ld], atol=1e-14)
    assert np.allclose(od.field("qacc"), [gen / 2, -2 * gen / 3], atol=1e-12)
    with pytest.raises(ox.MjsError, match="implicitfast"):
        ox.Model.from_xml_string(xml.replace('timestep="0.002"', 'timestep="0.002" integrator="implicitfast"'))


def test_compiler_refusals():
    body = "<worldbody><body><joint name='j' type='hinge'/><geom size='0.1'/><site name='s'/><body pos='0 0 1'><joint name='b' type='ball'/><geom size='0.1'/></body></body></worldbody>"
    for tendon, msg in (("<spatial><site site='s'/></spatial>", "two sites"), ("<spatial><site site='s'/><geom geom='g'/></spatial>", "outside the supported subset"), ("<fixed><joint joint='b' coef='1'/></fixed>", "hinge / slide"),
                        ("<fixed><joint joint='zz' coef='1'/></fixed>", "unknown joint"), ("<fixed frictionloss='1'><joint joint='j' coef='1'/></fixed>", "frictionloss"),
                        ("<fixed/>", "no joints")):
        with pytest.raises(ox.MjsError, match=msg):
            ox.Model.from_xml_string(f"<mujoco>{body}<tendon>{tendon}</tendon></mujoco>")
    with pytest.raises(ox.MjsError, match="implicitfast"):
        ox.Model.from_xml_string(f"<mujoco><option integrator='implicitfast'/>{body}<tendon><fixed damping='1'><joint joint='j' coef='1'/></fixed></tendon></mujoco>")


def test_host_instantiation_matches_oracle():
    m = ox.Model.from_xml_string(ZOO["zoo_h"])
    nenv, nsteps = 6, 200
    qpos, qvel = random_state(m, nenv, seed=43)
    qvel *= 10
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
        for f in ("qpos", "qvel", "qacc", "sensordata", "ten_length", "qfrc_passive"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-8, (f, e)
    assert rows > 100


@pytest.mark.gpu
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("coop", 0), ("fused", 2)])
def test_gpu_vs_oracle(mode, specialize):
    m = ox.Model.from_xml_string(ZOO["zoo_h"])
    nenv, nsteps = 64, 150
    qpos, qvel = random_state(m, nenv, seed=47)
    qvel *= 10
    if mode == "coop":   # the lane = dof kernel gathers actuator forces per joint: tendon transmissions are not offered there
        with pytest.raises(ox.Error, match="coop"):
            ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
        return
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    b.step(nsteps); b.sync()
    ref = {f: [] for f in ("qpos", "qacc", "sensordata")}
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        for f in ref:
            ref[f].append(od.field(f).copy())
    for f in ref:
        assert rel_err(b.get(f), np.stack(ref[f])) <= 1e-6, f
    assert int(b.diverged().sum()) == 0


SPATIAL = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 0"/><worldbody>
<site name="o" pos="0 0 0"/>
<body name="a" pos="0.3 0 0.4"><freejoint/><geom type="sphere" size="0.05" mass="2" contype="0" conaffinity="0"/><site name="sa" pos="0 0 0"/></body>
</worldbody><tendon><spatial name="t" {attrs}><site site="o"/><site site="sa"/></spatial></tendon>
<sensor><tendonpos tendon="t"/><tendonvel tendon="t"/></sensor></mujoco>"""


def test_spatial_tendon_closed_forms():
    """A point mass on a string from the origin: L = |p|, Ldot = p.v / |p|, the spring pulls along -p / |p| with k (L - L0), and the
    limit row is the unit radial direction (J = -side * p / |p| on the translational dofs)."""
    m = ox.Model.from_xml_string(SPATIAL.format(attrs='stiffness="30" damping="2" springlength="0.2"'))
    assert list(m.tendon_type) == [1] and abs(m.tendon_length0[0] - 0.5) < 1e-15
    assert abs(m.tendon_invweight0[0] - 0.5) < 1e-12                      # J M^-1 J' = 1/m for a unit radial Jacobian
    od = OracleData(m)
    p, v = np.array([0.2, -0.1, 0.3]), np.array([0.4, 0.2, -0.5])
    od.field("qpos")[:3] = p; od.field("qvel")[:3] = v
    od.forward()
    L, u = np.linalg.norm(p), p / np.linalg.norm(p)
    assert np.allclose(od.field("sensordata"), [L, u @ v], atol=1e-14) and np.allclose(od.field("ten_J")[:3], u, atol=1e-14)
    f = -30 * (L - 0.2) - 2 * (u @ v)
    assert np.allclose(od.field("qfrc_passive")[:3], f * u, atol=1e-13) and np.allclose(od.field("qacc")[:3], f * u / 2, atol=1e-12)
    lim = ox.Model.from_xml_string(SPATIAL.format(attrs='limited="true" range="0 0.45"'))
    od = OracleData(lim)
    od.forward()                                                          # L0 = 0.5 > 0.45: the upper limit is violated by 0.05
    assert od.int("nefc") == 1 and abs(od.field("efc_pos")[0] + 0.05) < 1e-15
    assert np.allclose(od.field("efc_J")[:3], -np.array([0.3, 0, 0.4]) / 0.5, atol=1e-14) and od.field("efc_force")[0] > 0
    for _ in range(2000):
        od.step()
    assert np.linalg.norm(od.field("qpos")[:3]) < 0.452                   # pulled back inside the limit


def test_spatial_tendon_host_instantiation():
    m = ox.Model.from_xml_string(ZOO["zoo_q"])
    nenv, nsteps = 5, 250
    qpos, qvel = random_state(m, nenv, seed=103)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        for f in ("qpos", "qvel", "qacc", "sensordata", "ten_length", "ten_J"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-8, (f, e)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("fused", 2)])
def test_spatial_tendon_gpu_vs_oracle(mode, specialize):
    m = ox.Model.from_xml_string(ZOO["zoo_q"])
    nenv, nsteps = 64, 150
    qpos, qvel = random_state(m, nenv, seed=107)
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    b.step(nsteps); b.sync()
    ref = {f: [] for f in ("qpos", "qacc", "sensordata")}
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        for f in ref:
            ref[f].append(od.field(f).copy())
    for f in ref:
        assert rel_err(b.get(f), np.stack(ref[f])) <= 1e-6, f
    assert int(b.diverged().sum()) == 0
