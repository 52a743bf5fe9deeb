"""N3 rows of SURVEY 8f: equality constraints (`eq_active`, /root/reference/src/physics.rs:147-152) and mocap bodies
(`mocap_pos` / `mocap_quat`, src/physics.rs:154-170).

CPU: closed forms pin the oracle (a single soft row on a one-dof system has the solution qacc = (1-d) a0 + d aref);
tests/test_golden.py pins it on the whole zoo_e model against the dense checker; the product's stage templates instantiated
on the host must match the oracle. GPU: every kernel family against the oracle with user-moved mocap bodies and toggled
constraints, the single-env accessors returning Some / None like the reference, snapshot / restore carrying the new state."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

HINGE = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81"/><worldbody><body pos="0 0 1">
<joint name="h" type="hinge" axis="0 1 0" damping="0.2"/><geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.03" contype="0" conaffinity="0"/>
</body></worldbody><equality><joint name="hold" joint1="h" polycoef="{c0} 0 0 0 0" solref="{tc} {dr}" solimp="{d0} {d1} {w} 0.5 2"/></equality></mujoco>"""

FREEBALL = """<mujoco><option timestep="0.002"/><worldbody><body name="target" mocap="true" pos="0.1 -0.2 1.5"/>
<body name="ball" pos="0.1 -0.2 1.5"><freejoint/><geom type="sphere" size="0.1" contype="0" conaffinity="0"/></body></worldbody>
<equality><connect name="c" body1="ball" body2="target" anchor="0 0 0"/></equality></mujoco>"""


def _impedance(d0, d1, width, x, mid=0.5, power=2.0):
    r = abs(x) / width
    if r >= 1:
        return d1
    y = r ** power / mid ** (power - 1) if r <= mid else 1 - (1 - r) ** power / (1 - mid) ** (power - 1)
    return d0 + y * (d1 - d0)


def test_joint_equality_closed_form_single_dof():
    """One hinge, one equality row: qacc = (1-d) a0 + d aref with aref = -b v - k d (q - c0), d = impedance(|q - c0|)."""
    c0, tc, dr, d0, d1, w = 0.3, 0.02, 1.0, 0.9, 0.95, 0.01
    m = ox.Model.from_xml_string(HINGE.format(c0=c0, tc=tc, dr=dr, d0=d0, d1=d1, w=w))
    assert (m.neq, m.nefcmax) == (1, 1) and int(m.eq_type[0]) == 2
    for q, v in ((0.3, 0.0), (0.304, 0.5), (0.1, -1.0), (0.29, 2.0)):
        od = OracleData(m)
        od.field("qpos")[0], od.field("qvel")[0] = q, v
        od.forward()
        a0 = float(od.field("qacc_smooth")[0])
        pos = q - c0
        d = _impedance(d0, d1, w, pos)
        k, b = 1 / (d1 * d1 * tc * tc * dr * dr), 2 / (d1 * tc)
        aref = -b * v - k * d * pos
        assert od.int("nefc") == 1 and od.int("ne") == 1
        assert abs(od.field("efc_aref")[0] - aref) <= 1e-12 * max(1, abs(aref))
        assert abs(od.field("qacc")[0] - ((1 - d) * a0 + d * aref)) <= 1e-10 * max(1, abs(aref)), (q, v)
    # the row pulls from BOTH sides (an inequality row would be inactive for Jaref > 0)
    od = OracleData(m)
    od.field("qpos")[0] = 0.1          # below the target: aref > 0 > a0, the row must still be active
    od.forward()
    assert od.field("efc_force")[0] > 0
    od.field("qpos")[0] = 0.5          # above: force of the other sign
    od.forward()
    assert od.field("efc_force")[0] < 0
    # switching it off removes the row
    od.field("eq_active")[0] = 0
    od.forward()
    assert od.int("nefc") == 0 and abs(od.field("qacc")[0] - od.field("qacc_smooth")[0]) == 0
    od.reset()
    assert od.field("eq_active")[0] == 1


def test_connect_to_mocap_closed_form_and_steady_state():
    """A free ball pinned at its centre to a mocap body: three decoupled rows with diagApprox = 1/m, so per axis
    qacc = (1-d) a0 + d aref; moving the mocap body drags the ball to the new position."""
    m = ox.Model.from_xml_string(FREEBALL)
    assert (m.nmocap, m.neq, m.nefcmax) == (1, 1, 3) and list(m.body_mocapid) == [-1, 0, -1]
    od = OracleData(m)
    assert np.allclose(od.field("mocap_pos"), [0.1, -0.2, 1.5]) and np.allclose(od.field("mocap_quat"), [1, 0, 0, 0])
    od.forward()
    d = 0.9                                                               # residual 0: impedance d0
    assert np.allclose(od.field("qacc")[:3], [0, 0, (1 - d) * -9.81], atol=1e-10)
    assert np.allclose(od.field("qacc")[3:], 0, atol=1e-12)
    target = np.array([0.15, -0.1, 1.45])
    od.field("mocap_pos")[:] = target
    od.field("mocap_quat")[:] = [2, 0, 0, 0]                              # un-normalised on purpose: mj_kinematics normalises the copy
    for _ in range(1500):
        od.step()
    sag = od.field("qpos")[:3] - target                                   # soft constraint: a small steady-state sag under gravity
    assert abs(sag[0]) < 1e-6 and abs(sag[1]) < 1e-6 and -2e-3 < sag[2] < 0
    assert np.allclose(od.field("xpos")[3:6], target) and np.allclose(od.field("xquat")[4:8], [1, 0, 0, 0])
    # steady state: constraint force balances gravity
    assert abs(od.field("qfrc_constraint")[2] - 9.81 * float(m.body_mass[2])) < 1e-4
    od.reset()
    assert np.allclose(od.field("mocap_pos"), [0.1, -0.2, 1.5])


def _inputs(m, nenv, seed):
    rng = np.random.default_rng(seed)
    qpos, qvel = random_state(m, nenv, seed=seed)
    od = OracleData(m)                 # keep it alive: field() returns views of its memory
    mpos = np.tile(od.field("mocap_pos"), (nenv, 1)) + rng.uniform(-0.05, 0.05, (nenv, 3 * m.nmocap))
    mquat = np.tile(od.field("mocap_quat"), (nenv, 1)) + rng.normal(0, 0.05, (nenv, 4 * m.nmocap))
    eqa = np.tile(np.asarray(m.eq_active0, dtype=np.float64), (nenv, 1))
    eqa[::3] = rng.integers(0, 2, eqa[::3].shape)
    return qpos, qvel, mpos, mquat, eqa


def _oracle(m, inputs, nsteps, nudge_at=None):
    qpos, qvel, mpos, mquat, eqa = inputs
    ods = []
    for e in range(qpos.shape[0]):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        od.field("mocap_pos")[:] = mpos[e]; od.field("mocap_quat")[:] = mquat[e]; od.field("eq_active")[:] = eqa[e]
        for s in range(nsteps):
            if s == nudge_at:
                od.field("mocap_pos")[:] += 0.03
            od.fill_ctrl_philox(e, s)
            od.step()
        ods.append(od)
    return ods


FIELDS = ["qpos", "qvel", "qacc", "sensordata", "qfrc_constraint", "xpos"]


def test_host_instantiation_matches_oracle_with_moving_mocap_and_toggled_constraints():
    m = ox.Model.from_xml_string(ZOO["zoo_e"])
    nenv, nsteps = 6, 100
    inputs = _inputs(m, nenv, 11)
    hb = HostBatch(m, nenv, "f64")
    for f, v in zip(("qpos", "qvel", "mocap_pos", "mocap_quat", "eq_active"), inputs):
        hb.set(f, v)
    hb.step(50, True, SEED, 0, 0)
    hb.set("mocap_pos", hb.get("mocap_pos") + 0.03)
    hb.step(nsteps - 50, True, SEED, 0, 50)
    ods = _oracle(m, inputs, nsteps, nudge_at=50)
    assert sum(od.int("ncon") for od in ods) > 0 and all(od.int("ne") >= 4 for od in ods[1:3])
    for f in FIELDS:
        assert rel_err(hb.get(f), np.stack([od.field(f) for od in ods])) <= 1e-7, f
    assert list(hb.get("nefc")[:, 0]) == [od.int("nefc") for od in ods]


def test_compiler_tables_and_refusals():
    m = ox.Model.from_xml_string(ZOO["zoo_e"])
    assert (m.neq, m.nmocap) == (5, 1)
    assert list(m.eq_type) == [0, 0, 2, 2, 0] and list(m.eq_active0) == [1, 1, 1, 0, 0]
    assert m.object_id(ox.obj.Equality, "couple").index == 2 and m.object_name(ox.ObjectId(ox.obj.Equality, 4)) == "pin"
    # the second anchor makes both anchors coincide at qpos0: fore tip (0.7, 0, 1) in the hand's frame
    hand = m.object_id(ox.obj.Body, "hand").index
    q = np.asarray(m.body_quat[4 * hand:4 * hand + 4]); q = q / np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    assert np.allclose(R @ np.asarray(m.eq_data[3:6]) + np.asarray(m.body_pos[3 * hand:3 * hand + 3]), [0.7, 0, 1.0], atol=1e-12)
    base = "<mujoco><worldbody><body name='a'><joint name='j' type='hinge'/><geom size='0.1'/><body name='b' pos='0 0 1'><joint name='k' type='ball'/><geom size='0.1'/></body></body></worldbody>{}</mujoco>"
    with pytest.raises(ox.MjsError, match="flex"):
        ox.Model.from_xml_string(base.format("<equality><flex flex='f'/></equality>"))
    with pytest.raises(ox.MjsError, match="hinge / slide"):
        ox.Model.from_xml_string(base.format("<equality><joint joint1='k'/></equality>"))
    with pytest.raises(ox.MjsError, match="unknown body"):
        ox.Model.from_xml_string(base.format("<equality><connect body1='zz' anchor='0 0 0'/></equality>"))
    with pytest.raises(ox.MjsError, match="child of the world"):
        ox.Model.from_xml_string("<mujoco><worldbody><body><geom size='0.1'/><body mocap='true'><geom size='0.1'/></body></body></worldbody></mujoco>")
    with pytest.raises(ox.MjsError, match="cannot have joints"):
        ox.Model.from_xml_string("<mujoco><worldbody><body mocap='true'><joint type='hinge'/><geom size='0.1'/></body></worldbody></mujoco>")


@pytest.mark.gpu
@pytest.mark.parametrize("mode,specialize", [("fused", False), ("staged", False), ("coop", False), ("fused", True)])
@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_gpu_vs_oracle(mode, specialize, precision):
    m = ox.Model.from_xml_string(ZOO["zoo_e"])
    nenv, nsteps = 64, 60
    inputs = _inputs(m, nenv, 13)
    b = ox.BatchedPhysics(m, nenv, precision=precision, mode=mode, specialize=2 if specialize else 0)
    if specialize:
        assert b.kernel_name().startswith("jit_"), b.kernel_name()
    for f, v in zip(("qpos", "qvel", "mocap_pos", "mocap_quat", "eq_active"), inputs):
        b.set(f, v)
    b.ctrl_philox(True, SEED)
    b.step(1); b.sync()
    ods = _oracle(m, inputs, 1)
    # fp32: the hook's solref 0.01 makes D ~ 1e5 / invweight, so cond(M + J'DJ) ~ 1e6 and eps * cond bounds the forward error of
    # qacc (|qacc| ~ 1e3 here); measured 1e-3 .. 3e-3 on qvel across the kernel families (tools/fp32_study.py has the analysis)
    tol1 = 1e-9 if precision == "f64" else 1e-2
    for f in ("qpos", "qvel", "qacc"):
        err = rel_err(b.get(f), np.stack([od.field(f) for od in ods]))
        print(mode, specialize, precision, f, "%.2e" % err)
        assert err <= tol1, f
    assert np.array_equal(b.get("nefc")[:, 0], [od.int("nefc") for od in ods])
    b.step(29); b.sync()
    b.set("mocap_pos", b.get("mocap_pos") + 0.03)
    b.step(nsteps - 30); b.sync()
    ods = _oracle(m, inputs, nsteps, nudge_at=30)
    assert sum(od.int("ncon") for od in ods) > 0
    tol = 1e-6 if precision == "f64" else 5e-2
    for f in ("qpos", "sensordata"):
        # fp32: the position sensors only - the force / torque sensors read the stiff hook's constraint force (|f| ~ 1e3), which two
        # fp32 trajectories a few 1e-2 apart do not share
        n = None if (precision == "f64" or f == "qpos") else 7
        err = rel_err(b.get(f)[:, :n], np.stack([od.field(f)[:n] for od in ods]))
        print(mode, specialize, precision, nsteps, "steps", f, "%.2e" % err)
        assert err <= tol, f
    assert int(b.diverged().sum()) == 0


@pytest.mark.gpu
def test_physics_accessors_and_snapshot():
    p = ox.Physics.from_xml_string(ZOO["zoo_e"])
    hand, ball = p.object_id(ox.obj.Body, "hand"), p.object_id(ox.obj.Body, "ball")
    hold = p.object_id(ox.obj.Equality, "hold")
    assert p.mocap_pos(ball) is None and p.set_mocap_pos(ball, [0, 0, 0]) is None and p.mocap_quat(ball) is None
    quat0 = np.array([0.98, 0, 0.2, 0]) / np.linalg.norm([0.98, 0.2])      # the compiler normalises body quaternions
    assert np.allclose(p.mocap_pos(hand), [0.72, 0.02, 0.95]) and np.allclose(p.mocap_quat(hand), quat0)
    assert p.eq_active(hold) is False and p.eq_active(p.object_id(ox.obj.Equality, "hook")) is True
    assert p.set_mocap_pos(hand, [0.7, 0.0, 0.9]) == () and p.set_mocap_quat(hand, [1, 0, 0, 0]) == ()
    p.set_eq_active(hold, True)
    od = OracleData(p.model())
    od.field("mocap_pos")[:] = [0.7, 0.0, 0.9]; od.field("mocap_quat")[:] = [1, 0, 0, 0]; od.field("eq_active")[3] = 1
    for _ in range(25):
        p.step(); od.step()
    assert rel_err(p.data().qpos, od.field("qpos")) <= 1e-9 and rel_err(p.data().xpos, od.field("xpos")) <= 1e-9
    p.reset()
    assert np.allclose(p.mocap_pos(hand), [0.72, 0.02, 0.95]) and p.eq_active(hold) is False
    # models without the features: None / error like the reference
    q = ox.Physics.from_xml_string(ox.models.PENDULUM)
    assert q.mocap_pos(ox.ObjectId(ox.obj.Body, 1)) is None
    # snapshot / restore carries mocap poses and eq_active
    m = ox.Model.from_xml_string(ZOO["zoo_e"])
    b = ox.BatchedPhysics(m, 16, precision="f64")
    inputs = _inputs(m, 16, 17)
    for f, v in zip(("qpos", "qvel", "mocap_pos", "mocap_quat", "eq_active"), inputs):
        b.set(f, v)
    b.ctrl_philox(True, SEED)
    b.step(10); b.sync()
    snap, counter = b.get_state(), b.step_counter()
    b.step(20); b.sync()
    ref = b.get("qpos")
    b.set("mocap_pos", np.zeros((16, 3))); b.set("eq_active", np.zeros((16, 5)))
    b.set_state(snap); b.set_step_counter(counter)
    b.step(20); b.sync()
    assert np.array_equal(b.get("qpos"), ref)


# ---------------------------------------------------------------------------------------------------- weld
WELDED = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 0"/><worldbody>
<body name="target" mocap="true" pos="0 0 1" quat="{q}"/>
<body name="b" pos="0 0 1"><freejoint/><geom type="box" size="0.1 0.2 0.3" contype="0" conaffinity="0"/></body></worldbody>
<equality><weld name="w" body1="b" body2="target" torquescale="{ts}"/></equality></mujoco>"""


def test_weld_rows_closed_form_and_tracking():
    """Free box welded to a mocap body. Rows 0-2 are the connect rows; rows 3-5 carry residual = torquescale * imag(conj(q2) q1 qrel)
    and J = G on the free joint's angular dofs (body axes for a free joint), so at small angles residual ~ torquescale * angle / 2."""
    ang = 0.2
    q = f"{np.cos(ang / 2)} 0 {np.sin(ang / 2)} 0"            # target rotated about y; qrel = that relative pose at qpos0
    m = ox.Model.from_xml_string(WELDED.format(q=q, ts=0.7))
    assert list(m.eq_type) == [1] and m.nefcmax == 6
    od = OracleData(m)
    od.forward()
    assert od.int("nefc") == 6 and od.int("ne") == 6 and np.allclose(od.field("efc_pos")[:6], 0, atol=1e-15)     # welded at qpos0
    od.field("mocap_quat")[:] = [1, 0, 0, 0]                  # move the target: the box must now turn by -ang about y
    od.forward()
    res = od.field("efc_pos")[3:6]
    # e = conj(q2) q1 qrel with q1 = q2 = identity: e = qrel, residual = ts * imag(qrel)
    assert np.allclose(res, 0.7 * np.array([0, np.sin(ang / 2), 0]), atol=1e-14)
    for _ in range(1500):
        od.step()
    qb = od.field("qpos")[3:7]
    assert np.allclose(qb * np.sign(qb[0]), [np.cos(ang / 2), 0, -np.sin(ang / 2), 0], atol=2e-3)   # orientation follows the mocap body
    assert np.abs(od.field("qpos")[:3] - [0, 0, 1]).max() < 1e-4


def test_weld_host_instantiation_and_refusals():
    m = ox.Model.from_xml_string(ZOO["zoo_o"])
    assert list(m.eq_type) == [1, 1, 1] and list(m.eq_active0) == [1, 1, 0]
    nenv, nsteps = 5, 150
    qpos, qvel = random_state(m, nenv, seed=89)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        assert od.int("ne") == 12
        for f in ("qpos", "qvel", "qacc", "sensordata"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-8, (f, e)
    with pytest.raises(ox.MjsError, match="tendon"):
        ox.Model.from_xml_string("<mujoco><worldbody><body name='a'><joint/><geom size='0.1'/></body></worldbody><equality><tendon tendon1='t'/></equality></mujoco>")


@pytest.mark.gpu
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("coop", 0), ("fused", 2)])
def test_weld_gpu_vs_oracle(mode, specialize):
    m = ox.Model.from_xml_string(ZOO["zoo_o"])
    nenv, nsteps = 64, 100
    rng = np.random.default_rng(97)
    qpos, qvel = random_state(m, nenv, seed=97)
    od0 = OracleData(m)
    mpos = np.tile(od0.field("mocap_pos"), (nenv, 1)) + rng.uniform(-0.03, 0.03, (nenv, 3))
    mquat = np.tile(od0.field("mocap_quat"), (nenv, 1)) + rng.normal(0, 0.05, (nenv, 4))
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    b.set("qpos", qpos); b.set("qvel", qvel); b.set("mocap_pos", mpos); b.set("mocap_quat", mquat); b.ctrl_philox(True, SEED)
    b.step(nsteps); b.sync()
    ref = {f: [] for f in ("qpos", "qacc", "sensordata")}
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]; od.field("mocap_pos")[:] = mpos[e]; od.field("mocap_quat")[:] = mquat[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        for f in ref:
            ref[f].append(od.field(f).copy())
    for f in ref:
        assert rel_err(b.get(f), np.stack(ref[f])) <= 1e-6, f
    assert int(b.diverged().sum()) == 0
