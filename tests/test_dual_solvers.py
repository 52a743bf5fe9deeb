"""N4 rows of SURVEY 8f: the PGS solver and the noslip post-pass (same entry as the step, /root/reference/src/physics.rs:44-46;
MuJoCo: mj_solPGS / mj_solNoSlip selected by <option solver="PGS" noslip_iterations="n"/>).

CPU: PGS must land on the minimiser the Newton solver finds (same convex problem, dual vs primal); a single-row problem
has the closed form f = max(0, -b / (A + R)); noslip must remove the creep that soft friction leaves on a slope.
tests/test_golden.py pins the matrix-free oracle against the dense checker's explicit-matrix PGS / noslip on zoo_f / zoo_g.
GPU: every kernel family that accepts these options against the oracle; the cooperative kernels refuse them loudly."""
import re

import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO


def with_option(xml, **opt):
    m = re.search(r"<option([^>]*?)(/?)>", xml)
    attrs = m.group(1) + "".join(f' {k}="{v}"' for k, v in opt.items())
    return xml[:m.start()] + "<option" + attrs + m.group(2) + ">" + xml[m.end():]


def test_pgs_converges_to_the_newton_minimiser():
    newton = ox.Model.from_xml_string(ox.models.CHEETAH)
    pgs = ox.Model.from_xml_string(with_option(ox.models.CHEETAH, solver="PGS", iterations=2000, tolerance="1e-15"))
    assert (newton.solver, pgs.solver) == (2, 0)
    qpos, qvel = random_state(newton, 6, seed=1)
    worst, rows = 0.0, 0
    for e in range(6):
        a, b = OracleData(newton), OracleData(pgs)
        a.field("qpos")[:] = qpos[e]; a.field("qvel")[:] = qvel[e]
        for s in range(150):
            a.fill_ctrl_philox(e, s); a.step()
        for f in ("qpos", "qvel", "qacc_warmstart"):
            b.field(f)[:] = a.field(f)
        for od in (a, b):
            od.fill_ctrl_philox(e, 999); od.forward()
        assert a.int("nefc") == b.int("nefc")
        rows += a.int("nefc")
        worst = max(worst, rel_err(b.field("qacc"), a.field("qacc")), rel_err(b.field("efc_force")[:a.int("nefc")], a.field("efc_force")[:a.int("nefc")]))
    assert rows > 20 and worst < 1e-7, worst


def test_pgs_single_row_closed_form():
    """Sphere resting on a frictionless plane: one row, A = 1/m, so f = max(0, -b / (1/m + R)) after ONE sweep."""
    xml = """<mujoco><option solver="PGS" iterations="1"><flag warmstart="disable"/></option><worldbody><geom type="plane" size="1 1 0.1" condim="1"/>
    <body pos="0 0 0.099"><freejoint/><geom type="sphere" size="0.1" condim="1"/></body></worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m)
    od.field("qvel")[2] = -0.05
    od.forward()
    assert od.int("nefc") == 1 and od.int("solver_niter") == 1
    mass, R, aref = float(m.body_mass[1]), 1 / od.field("efc_D")[0], od.field("efc_aref")[0]
    b = -9.81 - aref
    f = max(0.0, -b / (1 / mass + R))
    assert f > 0 and abs(od.field("efc_force")[0] - f) <= 1e-12 * f
    assert abs(od.field("qacc")[2] - (-9.81 + f / mass)) <= 1e-12


SLOPE = """<mujoco><compiler angle="radian"/><option timestep="0.002" {opt}/><worldbody>
<geom type="plane" size="3 3 0.1" euler="0 0.3 0" friction="1 0.005 0.0001"/>
<body pos="0 0 0.2" euler="0 0.3 0"><freejoint/><geom type="box" size="0.1 0.1 0.05" friction="1 0.005 0.0001"/></body></worldbody></mujoco>"""


def test_noslip_removes_friction_creep_on_a_slope():
    """tan(0.3) = 0.31 < mu = 1: the box should stick. Soft (regularised) friction lets it creep; noslip re-solves the friction
    dimensions without regularisation and holds it."""
    drift = {}
    for label, opt in (("soft", ""), ("noslip", 'noslip_iterations="10"'), ("pgs_noslip", 'solver="PGS" noslip_iterations="10"')):
        m = ox.Model.from_xml_string(SLOPE.format(opt=opt))
        od = OracleData(m)
        for _ in range(300):      # settle onto the slope
            od.step()
        p0 = od.field("qpos")[:3].copy()
        for _ in range(500):
            od.step()
        drift[label] = float(np.linalg.norm(od.field("qpos")[:3] - p0))
        assert od.int("ncon") == 4
    assert drift["soft"] > 1e-4 and drift["noslip"] < 0.05 * drift["soft"] and drift["pgs_noslip"] < 0.05 * drift["soft"], drift


@pytest.mark.parametrize("name", ["zoo_f", "zoo_g"])
def test_host_instantiation_matches_oracle(name):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 5, 150
    qpos, qvel = random_state(m, nenv, seed=23)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    niter = 0
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        niter += od.int("solver_niter")
        for f in ("qpos", "qvel", "qacc", "efc_force", "qfrc_constraint", "sensordata"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-8, (f, e)
        assert hb.get("solver_niter")[e, 0] == od.int("solver_niter")
    assert niter > 0


def test_options_are_compiled_and_cooperative_kernels_are_not_offered():
    m = ox.Model.from_xml_string(ZOO["zoo_f"])
    assert (m.solver, m.noslip_iterations, m.noslip_tolerance, m.iterations) == (0, 4, 1e-8, 60)
    assert ox.Model.from_xml_string(ox.models.CHEETAH).noslip_iterations == 0
    with pytest.raises(ox.MjsError, match="noslip_iterations"):
        ox.Model.from_xml_string(with_option(ox.models.CHEETAH, noslip_iterations=-1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zoo_f", "zoo_g"])
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("fused", 2)])
def test_gpu_vs_oracle(name, mode, specialize):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 64, 60
    qpos, qvel = random_state(m, nenv, seed=29)
    b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    if specialize:
        assert b.kernel_name().startswith("jit_"), b.kernel_name()
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
        assert np.array_equal(b.get("solver_niter")[:, 0], [od.int("solver_niter") for od in ods]) or upto > 1
    assert sum(od.int("ncon") for od in ods) > 0 and int(b.diverged().sum()) == 0


@pytest.mark.gpu
def test_cooperative_modes_refuse_dual_solvers():
    for name in ("zoo_f", "zoo_g"):
        m = ox.Model.from_xml_string(ZOO[name])
        with pytest.raises(ox.Error, match="coop"):
            ox.BatchedPhysics(m, 32, precision="f64", mode="coop")
        with pytest.raises(ox.Error, match="coop_solver"):
            ox.BatchedPhysics(m, 32, precision="f64", mode="staged", coop_solver=1)
