"""N3 row of SURVEY 8f: contact dimensions 4 and 6 - torsional friction (about the contact normal) and rolling friction (about the
tangents) as extra pyramid edge pairs on the relative ANGULAR velocity. Physical pins on the oracle (a spinning ball keeps its spin
with condim 3 and loses it with condim 4; a rolling ball coasts with condim 3/4 and stops with condim 6); tests/test_golden.py
pins zoo_j (oracle and CUDA) against the dense checker, including the torque part of the force / torque sensors."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

BALL = """<mujoco><option timestep="0.002"/><worldbody><geom type="plane" size="5 5 0.1" friction="1 0.05 0.02"/>
<body pos="0 0 0.0995"><freejoint/><geom type="sphere" size="0.1" condim="{dim}" friction="1 0.05 0.02"/></body></worldbody></mujoco>"""


def _run(dim, qvel, nsteps=600):
    m = ox.Model.from_xml_string(BALL.format(dim=dim))
    od = OracleData(m)
    for _ in range(200):
        od.step()                                    # settle on the floor
    od.field("qvel")[:] = qvel
    for _ in range(nsteps):
        od.step()
    assert od.int("ncon") == 1 and od.int("nefc") == (4 if dim == 3 else 2 * (dim - 1))
    return od.field("qvel").copy()


def test_torsional_friction_stops_a_spinning_ball_only_with_condim_4():
    spin = [0, 0, 0, 0, 0, 8.0]                      # about the contact normal
    assert abs(_run(3, spin)[5]) > 7.9               # a point contact cannot resist spin about its normal
    assert abs(_run(4, spin)[5]) < 0.5 and abs(_run(6, spin)[5]) < 0.5


def test_rolling_friction_stops_a_rolling_ball_only_with_condim_6():
    roll = [1.0, 0, 0, 0, 10.0, 0]                   # v = omega x r: pure rolling along +x
    v3, v4, v6 = _run(3, roll), _run(4, roll), _run(6, roll)
    assert v3[0] > 0.95 and v4[0] > 0.95             # rolling without slipping dissipates (almost) nothing
    assert abs(v6[0]) < 0.2 and abs(v6[4]) < 2.0     # rolling friction brings it to rest


def test_row_layout_and_compiler():
    m = ox.Model.from_xml_string(ZOO["zoo_j"])
    assert sorted(set(int(d) for d in m.pair_dim)) == [4, 6]
    with pytest.raises(ox.Error, match="condim"):
        ox.Model.from_xml_string(BALL.format(dim=5))
    od = OracleData(ox.Model.from_xml_string(BALL.format(dim=6)))
    od.field("qvel")[:] = [0.3, -0.2, 0, 1.0, 2.0, 3.0]
    od.forward()
    J = od.field("efc_J")[:60].reshape(10, 6)
    # edge pairs are J_n +/- mu_k J_k: their sum is twice the normal row, the same for every friction direction
    for k in range(5):
        assert np.allclose(J[2 * k] + J[2 * k + 1], J[0] + J[1], atol=1e-14)
    # torsion / rolling pairs differ only in the angular dofs (the ball centre is the free joint's origin... the contact point is
    # below it, so the tangential pairs also carry a lever arm on the angular dofs; the rotational pairs have NO linear part)
    for k in (2, 3, 4):
        assert np.allclose((J[2 * k] - J[2 * k + 1])[:3], 0, atol=1e-14) and np.abs((J[2 * k] - J[2 * k + 1])[3:]).max() > 1e-3


def test_host_instantiation_matches_oracle():
    m = ox.Model.from_xml_string(ZOO["zoo_j"])
    nenv, nsteps = 5, 250
    qpos, qvel = random_state(m, nenv, seed=61)
    qvel[:, 3:6] *= 40; qvel[:, 9:12] *= 40
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        assert od.int("nefc") >= 30
        for f in ("qpos", "qvel", "qacc", "sensordata", "efc_force"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-8, (f, e)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("coop", 0), ("fused", 2)])
def test_gpu_vs_oracle(mode, specialize):
    m = ox.Model.from_xml_string(ZOO["zoo_j"])
    nenv, nsteps = 64, 150
    qpos, qvel = random_state(m, nenv, seed=67)
    qvel[:, 3:6] *= 40; qvel[:, 9:12] *= 40
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        ods.append(od)
    done = 0
    for upto, tol in ((1, 1e-9), (nsteps, 1e-6)):
        b.step(upto - done); b.sync()
        for e, od in enumerate(ods):
            for s in range(done, upto):
                od.fill_ctrl_philox(e, s); od.step()
        done = upto
        for f in ("qpos", "qvel", "sensordata"):
            assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= tol, (f, upto)
    assert sum(od.int("nefc") for od in ods) > 30 * nenv and int(b.diverged().sum()) == 0
