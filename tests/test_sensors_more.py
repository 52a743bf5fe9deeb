"""N2 row of SURVEY 8f, second batch of sensor types read through Physics::data().sensordata (reference src/physics.rs:30-32):
framexaxis / frameyaxis / framezaxis, ballquat, ballangvel, jointactuatorfrc, framelinacc / frameangacc. Closed forms pin the oracle on zoo_a; the generic
zoo parity tests (tests/test_zoo_parity.py) compare every kernel family with the oracle on the same sensordata."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import OracleData, random_state
from zoo_models import ZOO
from tests_util import quat2mat


def _sensor(m, od, stype, k=0):
    idx = [i for i in range(m.nsensor) if int(m.sensor_type[i]) == stype][k]
    a, d = int(m.sensor_adr[idx]), int(m.sensor_dim[idx])
    return od.field("sensordata")[a:a + d], int(m.sensor_objid[idx])


def test_axes_ball_and_joint_actuator_force_closed_forms():
    m = ox.Model.from_xml_string(ZOO["zoo_a"])
    qpos, qvel = random_state(m, 3, seed=71)
    for e in range(3):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        od.fill_ctrl_philox(e, 0)
        od.forward()
        # frame axes are the columns of the object's rotation matrix
        x, sid = _sensor(m, od, 27)
        assert np.allclose(x, od.field("site_xmat").reshape(-1, 3, 3)[sid][:, 0], atol=1e-15)
        y, bid = _sensor(m, od, 28)
        assert np.allclose(y, od.field("xmat").reshape(-1, 3, 3)[bid][:, 1], atol=1e-15)
        z, gid = _sensor(m, od, 29)
        assert np.allclose(z, od.field("geom_xmat").reshape(-1, 3, 3)[gid][:, 2], atol=1e-15)
        assert abs(np.linalg.norm(z) - 1) < 1e-12
        # and agree with the framequat of the same kind of object: R(quat) e_x
        qs, qsid = _sensor(m, od, 26)
        assert np.allclose(quat2mat(qs), od.field("site_xmat").reshape(-1, 3, 3)[qsid], atol=1e-12)
        # ball joint: the joint's own (normalised) quaternion and angular velocity
        bq, jid = _sensor(m, od, 17)
        qa, da = int(m.jnt_qposadr[jid]), int(m.jnt_dofadr[jid])
        q = qpos[e][qa:qa + 4]
        assert np.allclose(bq, q / np.linalg.norm(q), atol=1e-15)
        bw, _ = _sensor(m, od, 18)
        assert np.array_equal(bw, qvel[e][da:da + 3])
        # net actuator force on a joint = sum over its actuators of gear * force
        for k in range(2):
            jf, j = _sensor(m, od, 16, k)
            want = sum(m.actuator_gear[a] * od.field("actuator_force")[a] for a in range(m.nu)
                       if int(m.actuator_trntype[a]) == 0 and int(m.actuator_trnid[a]) == j)
            assert abs(jf[0] - want) < 1e-14 and (k == 1 or abs(want) > 0)


def test_unnormalised_ball_quaternion_is_normalised_in_the_sensor():
    xml = """<mujoco><worldbody><body><joint name="b" type="ball"/><geom size="0.1"/></body></worldbody>
    <sensor><ballquat joint="b"/><ballangvel joint="b"/></sensor></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m)
    od.field("qpos")[:] = [2, 0, 0, 2]
    od.field("qvel")[:] = [0.1, 0.2, 0.3]
    od.forward()
    assert np.allclose(od.field("sensordata"), [np.sqrt(0.5), 0, 0, np.sqrt(0.5), 0.1, 0.2, 0.3], atol=1e-15)


def test_frame_accelerations():
    """framelinacc is the accelerometer's reading left in world axes (proper acceleration: a body at rest reads -gravity, in free
    fall 0); frameangacc of a hinged body is axis * qacc; and both are the time derivative of framelinvel / frameangvel."""
    m = ox.Model.from_xml_string(ZOO["zoo_a"])
    qpos, qvel = random_state(m, 3, seed=73)
    for e in range(3):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        od.fill_ctrl_philox(e, 0)
        od.forward()
        la, sid = _sensor(m, od, 32)
        acc = [(_sensor(m, od, 1, k)) for k in range(2)]
        same = [a for a, i in acc if i == sid][0]                       # zoo_a has an accelerometer on the same site
        assert np.allclose(od.field("site_xmat").reshape(-1, 3, 3)[sid] @ same, la, atol=1e-12)
    xml = """<mujoco><option timestep="1e-6" gravity="0 0 -9.81"/><worldbody>
    <body name="arm" pos="0 0 1"><joint name="h" axis="0 1 0" damping="0.1"/><geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.03"/>
      <site name="tip" pos="0.5 0 0.1"/>
      <body name="fore" pos="0.5 0 0"><joint name="k" axis="0.6 0 0.8"/><geom type="capsule" fromto="0 0 0 0.2 0.2 0" size="0.02"/><site name="end" pos="0.2 0.2 0"/></body>
    </body>
    <body name="drop" pos="1 0 1"><freejoint/><geom size="0.1"/></body></worldbody>
    <sensor><framelinacc objtype="site" objname="end"/><frameangacc objtype="site" objname="end"/><framelinvel objtype="site" objname="end"/>
    <frameangvel objtype="site" objname="end"/><frameangacc objtype="body" objname="arm"/><framelinacc objtype="body" objname="drop"/>
    <framelinacc objtype="xbody" objname="arm"/></sensor></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m)
    od.field("qpos")[:2] = [0.4, -0.7]; od.field("qvel")[:2] = [1.5, -2.0]
    od.field("qvel")[2:8] = [0.3, 0.1, -0.2, 1.0, 2.0, 3.0]
    od.forward()
    sd0 = od.field("sensordata").copy()
    assert np.allclose(sd0[12:15], np.array([0, 1, 0]) * od.field("qacc")[0], atol=1e-12)      # hinge about a fixed axis
    assert np.allclose(sd0[15:18], 0, atol=1e-12)                                              # free fall: no proper acceleration
    assert np.allclose(sd0[18:21], [0, 0, 9.81], atol=1e-12)                                   # the arm's origin is pinned: reads -gravity
    od.step()                                                                                   # sensors of the NEXT forward come with the next step
    od.forward()
    sd1 = od.field("sensordata")
    h = 1e-6
    assert np.allclose((sd1[6:9] - sd0[6:9]) / h, sd0[0:3] + np.array([0, 0, -9.81]), atol=2e-4)   # coordinate acceleration = proper + gravity
    assert np.allclose((sd1[9:12] - sd0[9:12]) / h, sd0[3:6], atol=2e-4)


def test_refusals():
    base = '<mujoco><worldbody><body><joint name="h"/><geom size="0.1"/></body></worldbody><sensor>{}</sensor></mujoco>'
    for s in ('<ballquat joint="h"/>', '<ballangvel joint="h"/>'):
        with pytest.raises(ox.MjsError, match="ball joint"):
            ox.Model.from_xml_string(base.format(s))
    ball = '<mujoco><worldbody><body><joint name="b" type="ball"/><geom size="0.1"/></body></worldbody><sensor>{}</sensor></mujoco>'
    with pytest.raises(ox.MjsError, match="hinge or slide"):
        ox.Model.from_xml_string(ball.format('<jointactuatorfrc joint="b"/>'))
