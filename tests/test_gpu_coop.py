"""mode = "coop" (csrc/ox_coop.cu): one lane group per environment for the whole step, intermediates in shared memory.
Generic (no model-specific code), so it is tested on every model it is eligible for: the BASELINE configs with Euler
integration, the user-supplied hopper and the zoo (ball / free joints, every contact primitive, sensors incl. touch and
accelerometer, stateful actuators, implicitfast). Gates are the usual ones: fp64 single step 1e-9 and 100-step horizon
against the oracle, the dense-checker fixtures, fp32 sanity."""
import json
import os

import numpy as np
import pytest

import oxide_control_b200 as ox
from support import OracleData, SEED, random_state, rel_err
from zoo_models import HOPPER, ZOO

pytestmark = pytest.mark.gpu
MODELS = {"pendulum": ox.models.PENDULUM, "cartpole": ox.models.CARTPOLE, "cheetah": ox.models.CHEETAH, "humanoid": ox.models.HUMANOID,
          "hopper": HOPPER, "zoo_a": ZOO["zoo_a"], "zoo_c": ZOO["zoo_c"]}


def _oracle(m, qpos, qvel, act, nsteps):
    ods = []
    for e in range(qpos.shape[0]):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        if m.na:
            od.field("act")[:] = act[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        ods.append(od)
    return ods


@pytest.mark.parametrize("name", list(MODELS))
def test_coop_step_matches_oracle_fp64(name):
    m = ox.Model.from_xml_string(MODELS[name])
    nenv, nsteps = 70, 100          # not a multiple of the groups per CTA: exercises the tail CTA
    qpos, qvel = random_state(m, nenv, seed=81)
    act = np.random.default_rng(81).uniform(-0.3, 0.3, (nenv, m.na))
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode="coop")
    assert "k_step_coop" in b.kernel_name()
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    if m.na:
        b.set("act", act)
    b.step(1); b.sync()
    ods = _oracle(m, qpos, qvel, act, 1)
    for f in ("qpos", "qvel", "qacc", "qacc_warmstart", "sensordata", "ctrl", "act", "time"):
        if b.field_size(f):
            assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= 1e-9, f
    assert np.array_equal(b.get("ncon")[:, 0], [od.int("ncon") for od in ods])
    assert np.array_equal(b.get("nefc")[:, 0], [od.int("nefc") for od in ods])
    b.step(nsteps - 1); b.sync()
    ods = _oracle(m, qpos, qvel, act, nsteps)
    assert np.max(np.abs(b.get("qpos") - np.stack([od.field("qpos") for od in ods]))) <= 1e-6
    assert rel_err(b.get("sensordata"), np.stack([od.field("sensordata") for od in ods])) <= 1e-5
    if m.npair:
        assert sum(od.int("ncon") for od in ods) > 0
    assert int(b.diverged().sum()) == 0
    with pytest.raises(ox.Error, match="not maintained"):
        b.get("xpos")                 # intermediates live in shared memory; forward() refreshes the batch copy
    b.forward(); b.sync()
    assert b.get("xpos").shape == (nenv, 3 * m.nbody)


@pytest.mark.parametrize("name", ["cheetah", "humanoid", "zoo_a", "zoo_c", "hopper"])
def test_coop_step_matches_dense_checker_fixtures(name):
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"dense_{name}.json")))
    m = ox.Model.from_xml_string(MODELS[name])
    cases = g["cases"]
    b = ox.BatchedPhysics(m, len(cases), precision="f64", mode="coop")
    for f in ("qpos", "qvel", "ctrl", "qfrc_applied", "xfrc_applied", "act"):
        v = np.array([c["input"][f] for c in cases], dtype=np.float64)
        if v.shape[1]:
            b.set(f, v)
    b.step(1); b.sync()
    assert np.array_equal(b.get("nefc")[:, 0], [c["output"]["nefc"] for c in cases])
    for f, t in (("qpos", 1e-9), ("qvel", 1e-9), ("qacc", 1e-8)):
        assert rel_err(b.get(f), np.array([c["output"][f] for c in cases])) <= t, f


@pytest.mark.parametrize("name", ["cheetah", "humanoid"])
def test_coop_fp32_and_multi_step_launch(name):
    m = ox.Model.from_xml_string(MODELS[name])
    nenv = 256
    qpos, qvel = random_state(m, nenv, seed=82)
    a = ox.BatchedPhysics(m, nenv, precision="f32", mode="coop")
    c = ox.BatchedPhysics(m, nenv, precision="f32", mode="coop")
    for x in (a, c):
        x.set("qpos", qpos); x.set("qvel", qvel); x.ctrl_philox(True, SEED)
    a.step(40)
    for _ in range(40):
        c.step(1)
    a.sync(); c.sync()
    assert np.array_equal(a.get("qpos"), c.get("qpos"))          # 40 steps in one launch == 40 launches, bit for bit
    ref = ox.BatchedPhysics(m, nenv, precision="f32")             # thread-per-env kernels, same precision
    ref.set("qpos", qpos); ref.set("qvel", qvel); ref.ctrl_philox(True, SEED); ref.step(1); ref.sync()
    one = ox.BatchedPhysics(m, nenv, precision="f32", mode="coop")
    one.set("qpos", qpos); one.set("qvel", qvel); one.ctrl_philox(True, SEED); one.step(1); one.sync()
    assert rel_err(one.get("qpos"), ref.get("qpos")) <= 1e-5 and np.isfinite(a.get("qpos")).all()


def test_coop_refuses_ineligible_models():
    with pytest.raises(ox.Error, match="coop"):
        ox.BatchedPhysics(ox.Model.from_xml_string(ox.models.ACROBOT), 8, mode="coop")      # RK4
    with pytest.raises(ox.Error, match="coop"):
        ox.BatchedPhysics(ox.Model.from_xml_string(ZOO["zoo_b"]), 8, mode="coop")           # CG solver
