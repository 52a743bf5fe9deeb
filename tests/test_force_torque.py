"""N2 remainder (VERDICT r1 missing #4): `force` and `torque` sensors - the wrench a body exchanges with its parent, from
mj_rnePostConstraint's cfrc_int (reads go through Physics::data(), /root/reference/src/physics.rs:30-32).

Closed forms pin the oracle; tests/test_golden.py checks it (and the CUDA path) on zoo_a / zoo_e against the dense checker's
independent momentum-balance derivation with contacts, applied Cartesian forces and connect constraints in play."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

ROD = """<mujoco><compiler angle="radian"/><option timestep="0.002"/><worldbody><body name="rod" pos="0 0 1">
<joint name="h" type="{jt}" axis="0 1 0"/><geom type="capsule" fromto="0 0 0 0.6 0 0" size="0.03" contype="0" conaffinity="0"/>
<site name="pivot" pos="0 0 0"/><site name="mid" pos="0.3 0 0" euler="0 0 1.5707963267948966"/></body></worldbody>
<actuator><motor joint="h"/></actuator>
<sensor><force site="pivot"/><torque site="pivot"/><force site="mid"/><torque site="mid"/></sensor></mujoco>"""


def test_horizontal_rod_released_from_rest_closed_form():
    """Rod pivoting about y, horizontal, at rest: qacc = +m g l / I_o (a +y rotation takes +x towards -z), the pivot force is
    m (a_com - g) = m g (1 - m l^2 / I_o) along +z, and the torque about the hinge axis equals the motor torque."""
    m = ox.Model.from_xml_string(ROD.format(jt="hinge"))
    od = OracleData(m)
    u = 0.35
    od.field("ctrl")[0] = u
    od.forward()
    mass, l = float(m.body_mass[1]), float(m.body_ipos[3])
    Io = float(od.field("qM")[0])
    qacc = (mass * 9.81 * l + u) / Io
    assert abs(od.field("qacc")[0] - qacc) < 1e-10
    sd = od.field("sensordata")
    fz = mass * (-qacc * l + 9.81)
    assert np.allclose(sd[0:3], [0, 0, fz], atol=1e-10)                  # force at the pivot, world-aligned site
    assert abs(sd[4] - u) < 1e-10 and abs(sd[3]) < 1e-12 and abs(sd[5]) < 1e-12   # torque through the joint = what the motor applies
    # the mid site is rotated 90 deg about z (x_site = y_world, y_site = -x_world): same force, re-expressed
    assert np.allclose(sd[6:9], [0, 0, fz], atol=1e-10)
    # torque about the mid site: pivot torque moved by r = (0.3, 0, 0): tau_mid = tau_pivot - r x F = (0, u + 0.3 fz, 0) -> site frame
    assert np.allclose(sd[9:12], [u + 0.3 * fz, 0, 0], atol=1e-10)


def test_free_fall_reads_zero_and_rest_reads_weight():
    xml = """<mujoco><worldbody><geom type="plane" size="1 1 0.1"/><body name="b" pos="0 0 {z}"><freejoint/>
    <geom type="sphere" size="0.1"/><site name="s"/></body></worldbody><sensor><force site="s"/><torque site="s"/></sensor></mujoco>"""
    od = OracleData(ox.Model.from_xml_string(xml.format(z=1.0)))
    od.forward()
    assert np.allclose(od.field("sensordata"), 0, atol=1e-12)             # free fall: nothing passes through the (free) joint
    m = ox.Model.from_xml_string(xml.format(z=0.0995))
    od = OracleData(m)
    for _ in range(2000):
        od.step()
    # resting on the floor: the contact force is external to the body, so the joint still carries nothing
    assert od.int("ncon") == 1 and np.allclose(od.field("sensordata"), 0, atol=1e-6)


@pytest.mark.parametrize("name", ["zoo_a", "zoo_e"])
def test_host_instantiation_matches_oracle(name):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 5, 90
    qpos, qvel = random_state(m, nenv, seed=31)
    xfrc = np.random.default_rng(31).normal(0, 1.5, (nenv, 6 * m.nbody)); xfrc[:, :6] = 0
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel); hb.set("xfrc_applied", xfrc)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]; od.field("xfrc_applied")[:] = xfrc[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        assert rel_err(hb.get("sensordata")[e], od.field("sensordata")) <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("coop", 0), ("fused", 2)])
def test_gpu_vs_oracle(mode, specialize):
    m = ox.Model.from_xml_string(ZOO["zoo_a"])
    nenv, nsteps = 64, 50
    qpos, qvel = random_state(m, nenv, seed=37)
    xfrc = np.random.default_rng(37).normal(0, 1.5, (nenv, 6 * m.nbody)); xfrc[:, :6] = 0
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    b.set("qpos", qpos); b.set("qvel", qvel); b.set("xfrc_applied", xfrc); b.ctrl_philox(True, SEED)
    b.step(nsteps); b.sync()
    ref, ncon = [], 0
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]; od.field("xfrc_applied")[:] = xfrc[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        ref.append(od.field("sensordata").copy()); ncon += od.int("ncon")
    assert ncon > 0 and rel_err(b.get("sensordata"), np.stack(ref)) <= 1e-6
