"""N2: the touch sensor (mjSENS_TOUCH), the last sensor type SURVEY 8f names. CPU: closed form pins the oracle (a resting
ball presses with its weight; a site that does not contain the contact point reads zero), and the product's templates on
the host match the oracle on a model with sphere / capsule / box sensing volumes. GPU: same through the C ABI."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

BALL = """<mujoco><compiler angle="radian"/><option timestep="0.002"/><worldbody>
<geom type="plane" size="2 2 0.1"/>
<body name="ball" pos="0 0 0.0995"><freejoint/><geom name="g" type="sphere" size="0.1" mass="2.5"/>
<site name="skin" type="sphere" size="0.12"/><site name="core" type="sphere" size="0.02"/>
<site name="cap_top" type="sphere" pos="0 0 0.1" size="0.05"/><site name="sole" type="box" pos="0 0 -0.1" size="0.05 0.05 0.02"/></body>
</worldbody><sensor><touch site="skin"/><touch site="core"/><touch site="cap_top"/><touch site="sole"/></sensor></mujoco>"""


def test_resting_ball_presses_with_its_weight():
    m = ox.Model.from_xml_string(BALL)
    od = OracleData(m)
    for _ in range(1500):
        od.step()
    od.forward()
    skin, core, top, sole = od.field("sensordata")
    assert od.int("ncon") == 1 and abs(od.field("qvel")).max() < 1e-6
    assert abs(skin - 2.5 * 9.81) < 1e-6 * 2.5 * 9.81          # normal force at rest = weight
    assert abs(sole - skin) < 1e-12                             # the box under the ball contains the contact point too
    assert core == 0.0 and top == 0.0                           # volumes the outward ray from the contact point never meets


def test_compiler_refuses_unsupported_touch_volumes():
    with pytest.raises(ox.MjsError, match="touch"):
        ox.Model.from_xml_string(BALL.replace('<site name="core" type="sphere" size="0.02"/>', '<site name="core" type="cylinder" size="0.02 0.02"/>'))


def test_host_instantiation_matches_oracle_on_touch_sensors():
    m = ox.Model.from_xml_string(ZOO["zoo_a"])
    nenv, nsteps = 8, 150
    qpos, qvel = random_state(m, nenv, seed=51)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    ref = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        ref.append(od.field("sensordata").copy())
    ref = np.array(ref)
    touch = ref[:, [int(a) for a, t in zip(m.sensor_adr, m.sensor_type) if t == 0]]      # the five touch sensors of zoo_a
    assert (touch[:, :4] > 0).sum() >= nenv and np.all(touch >= 0)       # ball, rod and box really are pressing on the floor
    assert rel_err(hb.get("sensordata"), ref) <= 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fused", "staged"])
def test_gpu_touch_sensors_match_oracle(mode):
    m = ox.Model.from_xml_string(ZOO["zoo_a"])
    nenv, nsteps = 64, 120
    qpos, qvel = random_state(m, nenv, seed=52)
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode)
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    b.step(nsteps); b.sync()
    ref = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        ref.append(od.field("sensordata").copy())
    ref = np.array(ref)
    assert (ref[:, [int(a) for a, t in zip(m.sensor_adr, m.sensor_type) if t == 0][:4]] > 0).sum() >= nenv
    assert rel_err(b.get("sensordata"), ref) <= 1e-6


@pytest.mark.gpu
def test_gpu_touch_resting_ball_specialised_at_run_time():
    m = ox.Model.from_xml_string(BALL)
    b = ox.BatchedPhysics(m, 32, precision="f64", specialize=2)
    assert b.kernel_name().startswith("jit_"), b.jit_note()
    b.step(1500); b.sync()
    s = b.get("sensordata")
    assert np.allclose(s[:, 0], 2.5 * 9.81, rtol=1e-6) and np.allclose(s[:, 3], s[:, 0]) and not s[:, 1:3].any()
