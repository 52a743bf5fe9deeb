"""N3 row of SURVEY 8f: dry joint friction (`frictionloss`): one Huber-cost row per dof - quadratic inside |J a - aref| < R f, force
pinned at +/- f outside - placed after the equality rows; a box constraint in PGS / noslip. Closed forms pin the oracle;
tests/test_golden.py pins zoo_k (Newton) and zoo_l (PGS + noslip) against the dense checker; GPU: kernel families vs the oracle
(the warp-cooperative kernels are not offered for models with friction rows)."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

HINGE = """<mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 0" {opt}/><worldbody><body>
<joint name="h" type="hinge" axis="0 1 0" frictionloss="0.3" damping="0.1"/><geom type="capsule" fromto="0 0 0 0.5 0 0" size="0.03" contype="0" conaffinity="0"/>
</body></worldbody><actuator><motor joint="h"/></actuator></mujoco>"""


@pytest.mark.parametrize("opt", ["", 'solver="CG"', 'solver="PGS"'])
def test_sliding_and_sticking_closed_forms(opt):
    m = ox.Model.from_xml_string(HINGE.format(opt=opt))
    assert m.nfloss == 1 and m.nefcmax == 1
    od = OracleData(m)
    I = None
    # moving fast: the friction row sits in its linear zone, the force is exactly -sign(v) * frictionloss
    for v, u in ((2.0, 0.5), (-3.0, 0.1), (1.5, -0.4)):
        od.field("qvel")[0], od.field("ctrl")[0] = v, u
        od.forward()
        I = float(od.field("qM")[0])
        assert od.int("nefc") == 1 and od.int("nf") == 1
        assert abs(od.field("efc_force")[0] + np.sign(v) * 0.3) < 1e-9
        assert abs(od.field("qacc")[0] - (u - 0.1 * v - np.sign(v) * 0.3) / I) < 1e-8
    # at rest with a torque below the friction level the joint stays put; above it, it breaks away
    for u, stuck in ((0.25, True), (-0.29, True), (0.45, False)):
        od = OracleData(m)
        od.field("ctrl")[0] = u
        for _ in range(500):
            od.step()
        assert (abs(od.field("qpos")[0]) < 5e-3) == stuck, (u, od.field("qpos")[0])
        if stuck:
            assert abs(od.field("efc_force")[0] + u) < 1e-3        # friction balances the applied torque
    # <flag frictionloss="disable"/> removes the row
    off = ox.Model.from_xml_string(HINGE.format(opt=opt).replace("/><worldbody>", '><flag frictionloss="disable"/></option><worldbody>', 1))
    od = OracleData(off)
    od.field("qvel")[0] = 1.0
    od.forward()
    assert od.int("nefc") == 0


@pytest.mark.parametrize("name", ["zoo_k", "zoo_l"])
def test_host_instantiation_matches_oracle(name):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 5, 250
    qpos, qvel = random_state(m, nenv, seed=71)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        assert od.int("nf") == 6
        for f in ("qpos", "qvel", "qacc", "sensordata", "qfrc_constraint"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-8, (f, e)
        assert rel_err(hb.get("efc_force")[e][:od.int("nefc")], od.field("efc_force")[:od.int("nefc")]) <= 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zoo_k", "zoo_l"])
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("fused", 2)])
def test_gpu_vs_oracle(name, mode, specialize):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 64, 150
    qpos, qvel = random_state(m, nenv, seed=73)
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
        for f in ("qpos", "qvel", "qacc"):
            assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= tol, (f, upto)
    assert int(b.diverged().sum()) == 0
    with pytest.raises(ox.Error, match="coop"):
        ox.BatchedPhysics(m, 32, precision="f64", mode="coop")
