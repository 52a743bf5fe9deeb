"""Parity on feature-coverage models (tests/zoo_models.py): ball/free/slide/hinge joints, box/sphere/capsule vs plane,
sphere-sphere, sphere-capsule and capsule-capsule pairs, priorities / solmix / margin+gap / condim 1, excludes, position /
velocity / affine general actuators with force clamps, every supported sensor, CG and Newton, Euler and RK4 with contacts,
applied Cartesian and joint forces, warm start disabled. CPU: the kernels' stage templates instantiated on the host vs the
oracle; GPU (marked): the generic CUDA kernels through the C ABI vs the oracle."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

FIELDS = ["qpos", "qvel", "qacc", "sensordata", "qfrc_constraint", "actuator_force", "qfrc_passive", "site_xpos", "cvel"]


# zoo_i: seven light free bodies - a 2 N m torque spins a 6 cm ball up to hundreds of rad/s within the test horizon, and tumbling
# contacts at that speed amplify the last bit of round-off past any fixed gate; its applied forces and initial spins are scaled down
GENTLE = {"zoo_i": 0.02, "zoo_m": 0.05, "zoo_n": 0.05, "zoo_p": 0.02}
# elliptic cones (zoo_m, zoo_n): the objective is not piecewise quadratic, so two implementations that stop on the same
# tolerance rule can sit a solver-tolerance apart; 80 steps of contact dynamics amplify that past the 1e-7 used elsewhere
HORIZON_TOL = {"zoo_m": 1e-5, "zoo_n": 1e-5}


def _inputs(m, nenv, seed, gentle=1.0):
    rng = np.random.default_rng(seed)
    qpos, qvel = random_state(m, nenv, seed=seed)
    qvel *= 5 if gentle == 1.0 else 1
    xfrc = rng.normal(0, 2.0 * gentle, (nenv, 6 * m.nbody)); xfrc[:, :6] = 0
    xfrc[rng.random((nenv, 6 * m.nbody)) < 0.5] = 0
    qfrc = rng.normal(0, 0.5 * gentle, (nenv, m.nv))
    return qpos, qvel, xfrc, qfrc


def _oracle(m, qpos, qvel, xfrc, qfrc, nsteps):
    ods = []
    for e in range(qpos.shape[0]):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        od.field("xfrc_applied")[:] = xfrc[e]; od.field("qfrc_applied")[:] = qfrc[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s)
            od.step()
        ods.append(od)
    return ods


@pytest.mark.parametrize("name", list(ZOO))
def test_zoo_host_instantiation_vs_oracle(name):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 6, 80
    qpos, qvel, xfrc, qfrc = _inputs(m, nenv, 5, GENTLE.get(name, 1.0))
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel); hb.set("xfrc_applied", xfrc); hb.set("qfrc_applied", qfrc)
    hb.step(nsteps, True, SEED, 0, 0)
    ods = _oracle(m, qpos, qvel, xfrc, qfrc, nsteps)
    assert sum(od.int("ncon") for od in ods) > 0 and sum(od.int("nefc") for od in ods) > 0   # contacts really happen
    for f in FIELDS:
        ref = np.stack([od.field(f) for od in ods])
        assert rel_err(hb.get(f), ref) <= HORIZON_TOL.get(name, 1e-7), f           # 80 steps of contact dynamics amplify round-off
    assert list(hb.get("ncon")[:, 0]) == [od.int("ncon") for od in ods]
    assert int(hb.get("diverged").sum()) == 0


@pytest.mark.parametrize("name", list(ZOO))
def test_zoo_single_step_every_field_host(name):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv = 8
    qpos, qvel, xfrc, qfrc = _inputs(m, nenv, 6, GENTLE.get(name, 1.0))
    # drop everything 5 cm so that the first step already has contacts
    for j in range(m.njnt):
        if int(m.jnt_type[j]) == 0:
            qpos[:, int(m.jnt_qposadr[j]) + 2] -= 0.05
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel); hb.set("xfrc_applied", xfrc); hb.set("qfrc_applied", qfrc)
    hb.step(1, True, SEED, 0, 0)
    ods = _oracle(m, qpos, qvel, xfrc, qfrc, 1)
    assert sum(od.int("nefc") for od in ods) > 0
    for f in FIELDS + ["qfrc_smooth", "qacc_smooth", "subtree_com", "geom_xpos", "xquat"]:
        ref = np.stack([od.field(f) for od in ods])
        assert rel_err(hb.get(f), ref) <= 1e-9, f


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(ZOO))
@pytest.mark.parametrize("mode", ["fused", "staged"])
def test_zoo_gpu_vs_oracle(name, mode):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 64, 40
    qpos, qvel, xfrc, qfrc = _inputs(m, nenv, 7, GENTLE.get(name, 1.0))
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode)
    b.set("qpos", qpos); b.set("qvel", qvel); b.set("xfrc_applied", xfrc); b.set("qfrc_applied", qfrc)
    b.ctrl_philox(True, SEED)
    b.step(1); b.sync()
    ods1 = _oracle(m, qpos, qvel, xfrc, qfrc, 1)
    for f in FIELDS:
        ref = np.stack([od.field(f) for od in ods1])
        assert rel_err(b.get(f), ref) <= (1e-5 if name in HORIZON_TOL else 1e-9), f
    b.step(nsteps - 1); b.sync()
    ods = _oracle(m, qpos, qvel, xfrc, qfrc, nsteps)
    assert sum(od.int("ncon") for od in ods) > 0
    assert rel_err(b.get("qpos"), np.stack([od.field("qpos") for od in ods])) <= max(1e-6, 10 * HORIZON_TOL.get(name, 0))
    assert rel_err(b.get("sensordata"), np.stack([od.field("sensordata") for od in ods])) <= max(1e-5, 10 * HORIZON_TOL.get(name, 0))
    assert int(b.diverged().sum()) == 0
