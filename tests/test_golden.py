"""Golden fixtures from the dense whole-step checker (tests/golden/dense_*.json, made by tools/make_golden.py).

CPU: the oracle against the fixtures, and the generator re-run on a few states (the fixtures are reproducible).
GPU: the CUDA path (fp64 validation mode; model-specialised, run-time-specialised or generic kernel, whatever the default
config selects, plus the generic fused kernel) against the same fixtures - i.e. NOT through the oracle.
The dense checker shares no algorithm with either (tests/dense_checker.py header); it does share the compiled model tables
and the author's reading of MuJoCo's documentation, so parity with libmujoco itself stays unpinned (test_golden_hook.py)."""
import json
import os

import numpy as np
import pytest

import oxide_control_b200 as ox
from support import OracleData, rel_err
from zoo_models import HOPPER, NOCONTACT, ZOO

INPUTS = ("qpos", "qvel", "ctrl", "qfrc_applied", "xfrc_applied", "act", "mocap_pos", "mocap_quat", "eq_active", "qacc_warmstart")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
XML = {**{k: v["xml"] for k, v in ox.models.CONFIGS.items()}, **ZOO, **NOCONTACT, "hopper": HOPPER}
# single-step gates, rel = |a-b| / max(1,|b|). zoo_b runs the CG solver (tolerance 1e-10, RK4: four solves per step), whose
# iterates stop at the tolerance instead of landing on the minimiser the way Newton's do - its qacc gate is the solver's.
TOL = {name: dict(qpos=1e-9, qvel=1e-9, qacc=1e-8, act=1e-12) for name in XML}
TOL["zoo_b"] = dict(qpos=1e-8, qvel=1e-5, qacc=1e-3, act=1e-12)
# elliptic cones: the objective is not piecewise quadratic (cone zone), so Newton / CG stop at their tolerance short of the exact
# minimiser the autodiff checker iterates to
# zoo_p: stacked boxes, up to 13 contacts with 52 strongly coupled pyramid rows: the Newton iteration stops on its improvement rule
# a few 1e-7 (relative) from the minimiser
TOL["zoo_p"] = dict(qpos=1e-9, qvel=1e-8, qacc=1e-5, act=1e-12)
TOL["zoo_m"] = dict(qpos=1e-9, qvel=1e-8, qacc=1e-5, act=1e-12)
TOL["zoo_n"] = dict(qpos=1e-7, qvel=1e-5, qacc=1e-4, act=1e-12)   # CG at tolerance 1e-12 on a non-quadratic objective
# derived arrays of the last forward of the step. With RK4 that is the 4th stage, whose inputs carry the (CG-tolerance) error
# of the first three solves in zoo_b; everywhere else they are functions of the input state alone.
AUX = {name: 1e-9 for name in XML}
AUX["zoo_b"] = 1e-4


def load(name):
    with open(os.path.join(GOLDEN, f"dense_{name}.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", list(XML))
def test_oracle_matches_dense_checker_fixtures(name):
    g = load(name)
    m = ox.Model.from_xml_string(XML[name])
    assert (g["nq"], g["nv"]) == (m.nq, m.nv) and len(g["cases"]) >= (64 if m.npair else 16)
    if m.npair:
        assert g["states_in_contact"] >= 30      # the fixtures really exercise the constrained path
    worst = dict(qpos=0.0, qvel=0.0, qacc=0.0, act=0.0)
    for case in g["cases"]:
        i, o = case["input"], case["output"]
        od = OracleData(m)
        for f in INPUTS:
            if len(i.get(f, ())):
                od.field(f)[:] = i[f]
        od.step()
        assert od.int("ncon") == o["ncon"] and od.int("nefc") == o["nefc"]
        n = o["nefc"]
        aux = AUX[name]
        assert rel_err(np.sort(od.field("efc_D")[:n]), o["efc_D_sorted"]) <= aux           # row order may differ, the row set may not
        assert rel_err(np.sort(od.field("efc_aref")[:n]), o["efc_aref_sorted"]) <= aux
        assert rel_err(np.sort(od.field("con_dist")[:o["ncon"]]), o["con_dist_sorted"]) <= aux
        assert rel_err(od.field("qfrc_bias"), o["qfrc_bias"]) <= aux
        assert rel_err(od.field("qfrc_smooth"), o["qfrc_smooth"]) <= aux
        assert rel_err(od.field("actuator_force"), o["actuator_force"]) <= aux
        if "ft_adr" in o:    # force / torque sensors, derived independently (momentum balance of the subtree) by the dense checker
            got = np.concatenate([od.field("sensordata")[a:a + 3] for a in o["ft_adr"]])
            assert rel_err(got, o["ft_val"]) <= 1e-8, rel_err(got, o["ft_val"])
        for f in worst:
            worst[f] = max(worst[f], rel_err(od.field(f), o[f]))
    print(name, worst)
    for f, t in TOL[name].items():
        assert worst[f] <= t, (f, worst[f])


@pytest.mark.parametrize("name", ["cheetah", "humanoid", "zoo_a", "zoo_c", "zoo_e", "zoo_f"])
def test_fixtures_are_reproducible(name):
    import dense_checker as dc
    g = load(name)
    dm = dc.DenseModel(ox.Model.from_xml_string(XML[name]))
    for case in g["cases"][-2:]:
        i, o = case["input"], case["output"]
        r = dc.step(dm, np.array(i["qpos"]), np.array(i["qvel"]), np.array(i["ctrl"]), np.array(i["qfrc_applied"]), np.array(i["xfrc_applied"]),
                    np.array(i["act"]), mocap=(np.array(i.get("mocap_pos", [])), np.array(i.get("mocap_quat", []))),
                    eq_active=np.array(i["eq_active"]) if "eq_active" in i else None,
                    warmstart=np.array(i["qacc_warmstart"]) if "qacc_warmstart" in i else None)
        assert r["ncon"] == o["ncon"] and r["nefc"] == o["nefc"]
        for f in ("qpos", "qvel", "act", "qacc", "qfrc_bias"):
            assert rel_err(r[f], o[f]) <= 1e-11, f


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(XML))
@pytest.mark.parametrize("kernel", ["default", "generic"])
def test_cuda_path_matches_dense_checker_fixtures(name, kernel):
    g = load(name)
    m = ox.Model.from_xml_string(XML[name])
    cases = g["cases"]
    n = len(cases)
    b = ox.BatchedPhysics(m, n, precision="f64", specialize=(kernel == "default"))
    for f in INPUTS:
        v = np.array([c["input"].get(f, []) for c in cases], dtype=np.float64)
        if v.shape[1]:
            b.set(f, v)
    b.step(1); b.sync()
    assert np.array_equal(b.get("ncon")[:, 0], [c["output"]["ncon"] for c in cases])
    assert np.array_equal(b.get("nefc")[:, 0], [c["output"]["nefc"] for c in cases])
    for f, t in TOL[name].items():
        ref = np.array([c["output"][f] for c in cases])
        assert rel_err(b.get(f), ref) <= t, (f, rel_err(b.get(f), ref))
    if "ft_adr" in cases[0]["output"]:
        sd = b.get("sensordata")
        got = np.stack([np.concatenate([sd[e, a:a + 3] for a in c["output"]["ft_adr"]]) for e, c in enumerate(cases)])
        assert rel_err(got, np.array([c["output"]["ft_val"] for c in cases])) <= 1e-8
    assert int(b.diverged().sum()) == 0
