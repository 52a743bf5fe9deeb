"""GPU tests of the entry points that round 1 only reached indirectly: ox_batch_forward (Physics::forward ->
mj_forward, /root/reference/src/physics.rs:48-50) against the oracle's forward on every config and zoo model; the
"which fields are current" contract of the model-specialised kernel; one kernel launch per step."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

pytestmark = pytest.mark.gpu

DERIVED = ["xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat", "site_xpos", "site_xmat",
           "subtree_com", "cinert", "cdof", "qM", "qLD", "cvel", "cdof_dot", "qfrc_bias", "qfrc_passive", "actuator_force",
           "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qacc", "qacc_warmstart", "qfrc_constraint", "sensordata",
           "efc_D", "efc_aref", "efc_pos", "efc_force", "con_dist", "con_pos", "con_frame"]
STATE = ["qpos", "qvel", "ctrl", "time", "qfrc_applied", "xfrc_applied"]
MODELS = {**{k: v["xml"] for k, v in ox.models.CONFIGS.items()}, **ZOO}


@pytest.mark.parametrize("name", list(MODELS))
@pytest.mark.parametrize("precision,tol", [("f64", 1e-9)])
def test_forward_every_derived_field(name, precision, tol):
    if name in ("zoo_m", "zoo_n"):   # elliptic cones: a non-quadratic objective, both sides stop a solver-tolerance short of its minimiser
        tol = 1e-6
    m = ox.Model.from_xml_string(MODELS[name])
    nenv = 64
    rng = np.random.default_rng(21)
    qpos, qvel = random_state(m, nenv, seed=21)
    for j in range(m.njnt):  # free bodies 5 cm lower: contacts in the very first forward
        if int(m.jnt_type[j]) == 0:
            qpos[:, int(m.jnt_qposadr[j]) + 2] -= 0.05
    ctrl = rng.uniform(-1, 1, (nenv, m.nu))
    qfrc = rng.normal(0, 0.5, (nenv, m.nv))
    xfrc = rng.normal(0, 1.0, (nenv, 6 * m.nbody)); xfrc[:, :6] = 0
    b = ox.BatchedPhysics(m, nenv, precision=precision)   # default config: forward must work on a specialised batch too
    b.set("qpos", qpos); b.set("qvel", qvel); b.set("qfrc_applied", qfrc); b.set("xfrc_applied", xfrc)
    if m.nu:
        b.set("ctrl", ctrl)
    before = {f: b.get(f) for f in STATE if b.field_size(f)}
    b.forward(); b.sync()
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        od.field("qfrc_applied")[:] = qfrc[e]; od.field("xfrc_applied")[:] = xfrc[e]
        if m.nu:
            od.field("ctrl")[:] = ctrl[e]
        od.forward()
        ods.append(od)
    assert np.array_equal(b.get("ncon")[:, 0], [od.int("ncon") for od in ods])
    assert np.array_equal(b.get("nefc")[:, 0], [od.int("nefc") for od in ods])
    for f in DERIVED:
        if not b.field_size(f):
            continue
        ref = np.stack([od.field(f) for od in ods])
        got = b.get(f)
        if f.startswith("efc_") or f.startswith("con_"):   # only the first nefc / ncon entries are defined
            per = got.shape[1] // max(1, (m.nefcmax if f.startswith("efc_") else m.nconmax))
            for e, od in enumerate(ods):
                n = (od.int("nefc") if f.startswith("efc_") else od.int("ncon")) * per
                assert rel_err(got[e, :n], ref[e, :n]) <= tol, (f, e)
        else:
            assert rel_err(got, ref) <= tol, f
    for f, v in before.items():   # mj_forward does not advance the state
        assert np.array_equal(b.get(f), v), f


def test_specialised_kernel_refuses_stale_derived_fields():
    """After a step of the model-specialised kernel only state / qacc / sensordata / counters are current; anything else
    must be refused until ox_batch_forward refreshes it - and then equal the generic kernels' values."""
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    nenv = 32
    qpos, qvel = random_state(m, nenv, seed=22)
    spec = ox.BatchedPhysics(m, nenv, precision="f64", specialize=True)
    gen = ox.BatchedPhysics(m, nenv, precision="f64", specialize=False)
    assert spec.kernel_name().startswith("cheetah") and "generic" in gen.kernel_name()
    for b in (spec, gen):
        b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED); b.step(3); b.sync()
    for f in ("qpos", "qvel", "qacc", "sensordata", "time", "ctrl"):
        assert rel_err(spec.get(f), gen.get(f)) <= 1e-11, f   # same algorithm; nvcc contracts FMAs differently in the two kernels
    with pytest.raises(ox.Error, match="not maintained by the model-specialised"):
        spec.get("xpos")
    with pytest.raises(ox.Error, match="not maintained by the model-specialised"):
        spec.get1("subtree_com", 0)
    assert gen.get("xpos").shape == (nenv, 3 * m.nbody)       # generic kernels keep everything current
    spec.forward(); gen.forward(); spec.sync(); gen.sync()
    for f in ("xpos", "xmat", "subtree_com", "cvel", "qfrc_bias", "qM"):
        assert rel_err(spec.get(f), gen.get(f)) <= 1e-10, f
    # an Environment task on a specialised batch may only observe / reward on maintained fields
    from oxide_control_b200.environment import BatchedEnvironment, TaskSpec
    with pytest.raises(ox.Error, match="not maintained by the model-specialised"):
        BatchedEnvironment(spec, TaskSpec(obs=[("xipos", 0, 3)]))


@pytest.mark.parametrize("name", ["cheetah", "cartpole", "acrobot"])
def test_one_launch_per_step(name):
    """gpu_launches == steps for the fused path, Philox controls included (the step counter is a kernel argument)."""
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    b = ox.BatchedPhysics(m, 256, precision="f32")
    b.ctrl_philox(True, SEED)
    n0 = b.launch_count()
    for _ in range(7):
        b.step(1)
    b.sync()
    assert b.launch_count() - n0 == 7


def test_step_counter_consistent_across_launch_shapes():
    """The Philox step index is the same whether steps go one per launch, many per launch, staged, or through a CUDA graph."""
    m = ox.Model.from_xml_string(ox.models.CARTPOLE)
    nenv = 64
    qpos, qvel = random_state(m, nenv, seed=23)
    outs = []
    for kw, chunks in ((dict(mode="fused"), [6]), (dict(mode="fused"), [1] * 6), (dict(mode="fused", specialize=False), [2, 4]),
                       (dict(mode="staged"), [3, 3]), (dict(mode="staged", use_graph=True), [1, 5])):
        b = ox.BatchedPhysics(m, nenv, precision="f64", **kw)
        b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
        for c in chunks:
            b.step(c)
        b.sync()
        outs.append((b.get("qpos"), b.get("ctrl")))
    for q, c in outs[1:]:
        assert rel_err(q, outs[0][0]) <= 1e-12 and np.array_equal(c, outs[0][1])
    # set_step_counter moves every path, the captured graph included
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode="staged", use_graph=True)
    a = ox.BatchedPhysics(m, nenv, precision="f64")
    for x in (a, b):
        x.set("qpos", qpos); x.set("qvel", qvel); x.ctrl_philox(True, SEED); x.step(2); x.set_step_counter(40); x.step(2); x.sync()
    assert np.array_equal(a.get("ctrl"), b.get("ctrl")) and rel_err(a.get("qpos"), b.get("qpos")) <= 1e-12
