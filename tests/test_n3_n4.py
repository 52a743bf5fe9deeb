"""N3 / N4 rows of SURVEY 8f: stateful actuators (`act`, /root/reference/src/physics.rs:96-102), the `implicitfast`
integrator and state snapshot / restore (same entry as the step, src/physics.rs:44-46).

CPU: closed forms pin the oracle (and, through tests/test_golden.py, the dense checker pins it on whole models);
the product's stage templates instantiated on the host must match the oracle. GPU: the CUDA path against the oracle,
snapshot -> restore -> identical trajectory, act accessors returning Some / None like the reference."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import NOCONTACT, ZOO

PEND = """<mujoco><compiler angle="radian"/><option timestep="0.01" integrator="{integ}" gravity="0 0 0"/><worldbody><body>
<joint name="h" type="hinge" axis="0 1 0" damping="0.3"/><geom type="capsule" fromto="0 0 0 0 0 -0.5" size="0.03" contype="0" conaffinity="0"/>
</body></worldbody><actuator>{act}</actuator></mujoco>"""


def _pend(integ, act):
    return ox.Model.from_xml_string(PEND.format(integ=integ, act=act))


def test_implicitfast_closed_form_velocity_servo():
    """One dof, no gravity: I qacc = -b w + kv (u - w). implicitfast solves (I + h (b + kv)) qacc' = I qacc; Euler only (I + h b)."""
    kv, b, u, w0 = 2.0, 0.3, 0.7, 1.3
    for integ, extra in (("implicitfast", kv), ("Euler", 0.0)):
        m = _pend(integ, f'<velocity joint="h" kv="{kv}"/>')
        od = OracleData(m)
        od.field("qvel")[0], od.field("ctrl")[0] = w0, u
        od.forward()
        I = float(od.field("qM")[0])
        qacc = (-b * w0 + kv * (u - w0)) / I
        assert abs(od.field("qacc")[0] - qacc) < 1e-12
        od.step()
        w1 = w0 + m.timestep * I * qacc / (I + m.timestep * (b + extra))
        assert abs(od.field("qvel")[0] - w1) < 1e-13, integ
        assert abs(od.field("qpos")[0] - m.timestep * w1) < 1e-14


def test_implicitfast_skips_force_clamped_actuators_and_is_more_stable_than_euler():
    m = _pend("implicitfast", '<velocity joint="h" kv="2" forcerange="-0.1 0.1"/>')
    od = OracleData(m)
    od.field("qvel")[0], od.field("ctrl")[0] = 1.3, -5.0     # force saturates at -0.1: no derivative term
    od.forward(); I = float(od.field("qM")[0]); qacc = float(od.field("qacc")[0])
    assert abs(od.field("actuator_force")[0] + 0.1) < 1e-15
    od.step()
    assert abs(od.field("qvel")[0] - (1.3 + m.timestep * I * qacc / (I + m.timestep * 0.3))) < 1e-13
    # a stiff velocity servo (kv h / I >> 2) explodes under Euler and decays under implicitfast
    for integ, stable in (("Euler", False), ("implicitfast", True)):
        m = _pend(integ, '<velocity joint="h" kv="400"/>')
        od = OracleData(m)
        od.field("qvel")[0] = 1.0
        for _ in range(5):
            od.step()
        assert (abs(od.field("qvel")[0]) < 1.0) == stable, (integ, od.field("qvel")[0])


def test_activation_dynamics_closed_forms():
    h, tau = 0.01, 0.04
    acts = ('<general name="i" joint="h" dyntype="integrator" actlimited="true" actrange="-0.05 0.05"/>'
            f'<general name="f" joint="h" dyntype="filter" dynprm="{tau}"/>'
            f'<general name="x" joint="h" dyntype="filterexact" dynprm="{tau}" gainprm="3"/>'
            '<motor name="m" joint="h"/>')
    m = _pend("Euler", acts)
    assert (m.na, m.nu) == (3, 4) and list(m.actuator_actadr) == [0, 1, 2, -1]
    od = OracleData(m)
    a0 = np.array([0.01, 0.2, -0.3]); u = np.array([0.9, 1.0, 0.5, 0.25])
    od.field("act")[:] = a0; od.field("ctrl")[:] = u
    od.forward()
    assert np.allclose(od.field("act_dot"), [u[0], (u[1] - a0[1]) / tau, (u[2] - a0[2]) / tau], atol=1e-14)
    assert np.allclose(od.field("actuator_force"), [a0[0], a0[1], 3 * a0[2], u[3]], atol=1e-15)   # the force comes from act, not ctrl
    od.step()
    exp = [min(0.05, a0[0] + h * u[0]), a0[1] + h * (u[1] - a0[1]) / tau, a0[2] + (u[2] - a0[2]) * (1 - np.exp(-h / tau))]
    assert np.allclose(od.field("act"), exp, atol=1e-15)
    for _ in range(20):
        od.step()
    assert od.field("act")[0] == 0.05                                                              # integrator pinned at actrange
    od.reset()
    assert not od.field("act").any()


@pytest.mark.parametrize("name", ["zoo_c", "zoo_d"])
def test_host_instantiation_matches_oracle(name):
    xml = {**ZOO, **NOCONTACT}[name]
    m = ox.Model.from_xml_string(xml)
    nenv, nsteps = 6, 120
    qpos, qvel = random_state(m, nenv, seed=41)
    act = np.random.default_rng(41).uniform(-0.3, 0.3, (nenv, m.na))
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel); hb.set("act", act)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]; od.field("act")[:] = act[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        for f in ("qpos", "qvel", "act", "act_dot", "qacc", "sensordata"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-8, (f, e)


def test_compiler_refusals_and_accessor_semantics_without_gpu():
    with pytest.raises(ox.MjsError, match="implicit"):
        _pend("implicit", '<motor joint="h"/>')
    with pytest.raises(ox.MjsError, match="dyntype"):
        _pend("Euler", '<general joint="h" dyntype="muscle"/>')
    with pytest.raises(ox.MjsError, match="actlimited"):
        _pend("Euler", '<general joint="h" actlimited="true" actrange="0 1"/>')
    assert _pend("implicitfast", '<motor joint="h"/>').integrator == 3


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zoo_c", "zoo_d"])
@pytest.mark.parametrize("kw", [dict(mode="fused", specialize=0), dict(mode="staged"), dict(mode="fused", specialize=2)])
def test_gpu_matches_oracle(name, kw):
    xml = {**ZOO, **NOCONTACT}[name]
    m = ox.Model.from_xml_string(xml)
    nenv, nsteps = 64, 60
    qpos, qvel = random_state(m, nenv, seed=42)
    act = np.random.default_rng(42).uniform(-0.3, 0.3, (nenv, m.na))
    b = ox.BatchedPhysics(m, nenv, precision="f64", **kw)
    if kw.get("specialize") == 2:
        assert b.kernel_name().startswith("jit_"), b.jit_note()
    b.set("qpos", qpos); b.set("qvel", qvel); b.set("act", act); b.ctrl_philox(True, SEED)
    b.step(1); b.sync()
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]; od.field("act")[:] = act[e]
        od.fill_ctrl_philox(e, 0); od.step()
        ods.append(od)
    for f in ("qpos", "qvel", "act", "act_dot", "qacc"):
        assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= 1e-9, f
    b.step(nsteps - 1); b.sync()
    for s in range(1, nsteps):
        for e, od in enumerate(ods):
            od.fill_ctrl_philox(e, s); od.step()
    for f in ("qpos", "act"):
        assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= 1e-6, f
    assert int(b.diverged().sum()) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,precision", [("cheetah", "f32"), ("cheetah", "f64"), ("zoo_c", "f64"), ("humanoid", "f32")])
def test_snapshot_restore_reproduces_the_trajectory(name, precision):
    xml = ox.models.CONFIGS[name]["xml"] if name in ox.models.CONFIGS else ZOO[name]
    m = ox.Model.from_xml_string(xml)
    nenv = 96
    qpos, qvel = random_state(m, nenv, seed=43)
    a = ox.BatchedPhysics(m, nenv, precision=precision)
    a.set("qpos", qpos); a.set("qvel", qvel); a.ctrl_philox(True, SEED)
    a.step(40); a.sync()
    snap, counter = a.get_state(np.float64), a.step_counter()
    assert snap.shape == (nenv, 1 + m.nq + 2 * m.nv + m.na + m.nu + m.nv + 6 * m.nbody) and counter == 40
    assert np.allclose(snap[:, 0], 40 * m.timestep) and np.array_equal(snap[:, 1:1 + m.nq], a.get("qpos"))
    a.step(25); a.sync()
    want_q, want_v, want_a = a.get("qpos"), a.get("qvel"), a.get("qacc")
    b = ox.BatchedPhysics(m, nenv, precision=precision)      # a fresh batch resumes from the checkpoint
    b.set_state(snap); b.ctrl_philox(True, SEED); b.set_step_counter(counter)
    b.step(25); b.sync()
    assert np.array_equal(b.get("qpos"), want_q) and np.array_equal(b.get("qvel"), want_v) and np.array_equal(b.get("qacc"), want_a)
    assert np.array_equal(b.get_state(np.float64), a.get_state(np.float64))


@pytest.mark.gpu
def test_act_accessors_mirror_the_reference():
    p = ox.Physics.from_xml_string(ZOO["zoo_c"])
    hip_int, knee_vel = p.object_id(ox.obj.Actuator, "hip_int"), p.object_id(ox.obj.Actuator, "knee_vel")
    assert p.act(knee_vel) is None and p.set_act(knee_vel, 1.0) is None          # stateless -> None (src/physics.rs:96-102)
    assert p.act(hip_int) == 0.0 and p.set_act(hip_int, 0.25) == ()
    assert p.act(hip_int) == 0.25
    p.set_ctrl(hip_int, 1.0)
    p.step()
    assert abs(p.act(hip_int) - (0.25 + 0.005)) < 1e-15                            # integrator: act += h ctrl
    p.reset()
    assert p.act(hip_int) == 0.0


def test_ball_joint_limit_closed_form():
    """A ball joint rotated by `ang` about a unit axis u, limit at 0.5 rad: one row, pos = 0.5 - ang, J = -u on the joint's dofs
    (mj_instantiateLimit); inside the limit there is no row."""
    xml = """<mujoco><compiler angle="radian"/><option gravity="0 0 0"/><worldbody><body pos="0 0 1">
    <joint name="b" type="ball" range="0 0.5" margin="0.01"/><geom type="capsule" fromto="0 0 0 0.4 0 0" size="0.03" contype="0" conaffinity="0"/>
    </body></worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    assert m.jnt_limited[0] == 1 and m.nefcmax >= 1
    u = np.array([1.0, 2.0, -2.0]) / 3.0
    for ang, rows in ((0.3, 0), (0.495, 1), (0.62, 1)):
        od = OracleData(m)
        od.field("qpos")[:] = np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * u])
        od.field("qvel")[:] = [0.1, -0.2, 0.3]
        od.forward()
        assert od.int("nefc") == rows, ang
        if rows:
            assert abs(od.field("efc_pos")[0] - (0.5 - ang)) < 1e-12 and np.allclose(od.field("efc_J")[:3], -u, atol=1e-12)
            assert od.field("efc_force")[0] > 0 if ang > 0.5 else True
    # the limit pushes back: released beyond the limit, the angle returns inside
    od = OracleData(m)
    od.field("qpos")[:] = np.concatenate([[np.cos(0.35)], np.sin(0.35) * u])      # 0.7 rad
    for _ in range(400):
        od.step()
    q = od.field("qpos")
    assert 2 * np.arctan2(np.linalg.norm(q[1:]), q[0]) < 0.52


def test_position_timeconst_is_a_filterexact_activation():
    """<position timeconst="T">: the target passes a first-order filter integrated exactly (dyntype filterexact, dynprm[0] = T);
    the force is kp (act - q) - kv v with the filtered target act."""
    xml = """<mujoco><option timestep="0.001" gravity="0 0 0"/><worldbody><body><joint name="j" type="slide" axis="1 0 0"/>
    <geom size="0.1" mass="1"/></body></worldbody><actuator><position joint="j" kp="10" kv="0.5" timeconst="0.05"/></actuator></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    assert m.na == 1 and int(m.actuator_dyntype[0]) == 3 and m.actuator_dynprm[0] == 0.05
    od = OracleData(m)
    od.field("ctrl")[:] = 1.0
    od.field("qpos")[:] = 0.2; od.field("qvel")[:] = 0.3
    od.forward()
    assert abs(od.field("actuator_force")[0] - (10 * (0.0 - 0.2) - 0.5 * 0.3)) < 1e-14        # act = 0 at the start
    for _ in range(100):
        od.step()
    assert abs(od.field("act")[0] - (1 - np.exp(-0.1 / 0.05))) < 1e-12                         # exact, not Euler
    od.forward()
    q, v, act = od.field("qpos")[0], od.field("qvel")[0], od.field("act")[0]
    assert abs(od.field("actuator_force")[0] - (10 * (act - q) - 0.5 * v)) < 1e-13
    with pytest.raises(ox.MjsError, match="dampratio"):
        ox.Model.from_xml_string(xml.replace('timeconst="0.05"', 'dampratio="1"').replace('kv="0.5"', ""))
