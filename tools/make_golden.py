#!/usr/bin/env python
"""Generates tests/golden/dense_<model>.json with the dense whole-step checker (tests/dense_checker.py).

    python tools/make_golden.py [model ...]

Each fixture holds N input states (qpos, qvel, ctrl, qfrc_applied, xfrc_applied) of one model and, for each, the outputs of ONE
mj_step computed by the dense checker - a dense numpy / torch-autodiff restatement that shares no algorithm with the oracle
or the CUDA kernels (see its header). The oracle is used here only to PRODUCE INTERESTING INPUT STATES (random initial state
rolled forward until the model is in contact); every stored output comes from the dense checker alone.
tests/test_golden.py checks the oracle (CPU) and the CUDA path (GPU) against these files, and re-runs the generator on a
few states to show the fixtures are reproducible. Real-MuJoCo vectors remain a separate hook (tools/dump_mujoco_golden.py).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oxide_control_b200 as ox  # noqa: E402
import dense_checker as dc  # noqa: E402
from support import OracleData, random_state  # noqa: E402
from zoo_models import HOPPER, NOCONTACT, ZOO  # noqa: E402

MODELS = {  # name -> (xml, number of states, apply external forces)
    "pendulum": (ox.models.PENDULUM, 16, False), "cartpole": (ox.models.CARTPOLE, 16, False), "acrobot": (ox.models.ACROBOT, 16, False),
    "cheetah": (ox.models.CHEETAH, 64, False), "humanoid": (ox.models.HUMANOID, 64, False),
    "zoo_a": (ZOO["zoo_a"], 64, True), "zoo_b": (ZOO["zoo_b"], 64, True), "hopper": (HOPPER, 64, True),
    "zoo_c": (ZOO["zoo_c"], 64, True), "zoo_d": (NOCONTACT["zoo_d"], 16, False), "zoo_e": (ZOO["zoo_e"], 64, True),
    "zoo_f": (ZOO["zoo_f"], 64, True), "zoo_g": (ZOO["zoo_g"], 64, True),
    "zoo_h": (ZOO["zoo_h"], 64, True), "zoo_i": (ZOO["zoo_i"], 64, True),
    "zoo_j": (ZOO["zoo_j"], 64, True), "zoo_k": (ZOO["zoo_k"], 64, True), "zoo_l": (ZOO["zoo_l"], 64, True),
    "zoo_m": (ZOO["zoo_m"], 64, True), "zoo_n": (ZOO["zoo_n"], 64, True),
    "zoo_o": (ZOO["zoo_o"], 64, True), "zoo_p": (ZOO["zoo_p"], 64, True),
    "zoo_q": (ZOO["zoo_q"], 64, True), "zoo_r": (ZOO["zoo_r"], 64, True),
}
OUT_KEYS = ("qpos", "qvel", "act", "qacc", "qfrc_bias", "qfrc_smooth", "qfrc_constraint", "actuator_force")


def input_states(model, n, forces, seed=20261018):
    """Random initial states rolled forward on the oracle for 20 + 5e steps (e = state index) with Philox controls."""
    rng = np.random.default_rng(seed)
    qpos, qvel = random_state(model, n, seed=seed % 1000)
    states = []
    for e in range(n):
        od = OracleData(model)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        if model.na:
            od.field("act")[:] = rng.uniform(-0.5, 0.5, model.na)
        if model.nmocap:                      # the user moves the mocap bodies: a random pose near the model's
            od.field("mocap_pos")[:] += rng.uniform(-0.05, 0.05, 3 * model.nmocap)
            od.field("mocap_quat")[:] += rng.normal(0, 0.05, 4 * model.nmocap)      # left un-normalised on purpose (mj_kinematics normalises)
        if model.neq and e % 3 == 1:          # ... and switches equality constraints on and off
            od.field("eq_active")[:] = rng.integers(0, 2, model.neq)
        for s in range(20 + 5 * e if model.nefcmax else e):
            od.fill_ctrl_philox(e, s)
            od.step()
        od.fill_ctrl_philox(e, 100000)
        xf = np.zeros(6 * model.nbody)
        qf = np.zeros(model.nv)
        if forces and e % 2:
            xf = rng.normal(0, 1.0, 6 * model.nbody); xf[:6] = 0
            qf = rng.normal(0, 0.3, model.nv)
        states.append(dict(qpos=od.field("qpos").copy(), qvel=od.field("qvel").copy(), ctrl=od.field("ctrl").copy(), qfrc_applied=qf, xfrc_applied=xf,
                           act=od.field("act").copy(), mocap_pos=od.field("mocap_pos").copy(), mocap_quat=od.field("mocap_quat").copy(),
                           eq_active=od.field("eq_active").copy(), qacc_warmstart=od.field("qacc_warmstart").copy()))
    return states


def dense_case(dm, st):
    r = dc.step(dm, st["qpos"], st["qvel"], st["ctrl"], st["qfrc_applied"], st["xfrc_applied"], st["act"],
                mocap=(st["mocap_pos"], st["mocap_quat"]), eq_active=st["eq_active"], warmstart=st["qacc_warmstart"])
    out = {k: r[k] for k in OUT_KEYS}
    out.update(ncon=int(r["ncon"]), nefc=int(r["nefc"]), efc_D_sorted=np.sort(r["efc_D"]), efc_aref_sorted=np.sort(r["efc_aref"]),
               con_dist_sorted=np.sort(r["con_dist"]))
    if r["ft_sensors"]:   # force / torque sensors: sensordata addresses and the dense checker's values (of the forward pass of the step)
        out.update(ft_adr=[int(a) for a in sorted(r["ft_sensors"])], ft_val=np.concatenate([r["ft_sensors"][a] for a in sorted(r["ft_sensors"])]))
    return out


def tolist(d):
    return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def main():
    names = sys.argv[1:] or list(MODELS)
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name in names:
        xml, n, forces = MODELS[name]
        model = ox.Model.from_xml_string(xml)
        dm = dc.DenseModel(model)
        cases = []
        for st in input_states(model, n, forces):
            cases.append(dict(input=tolist(st), output=tolist(dense_case(dm, st))))
        ncon = [c["output"]["ncon"] for c in cases]
        doc = dict(model=name, generator="tools/make_golden.py (tests/dense_checker.py)", nq=model.nq, nv=model.nv, mean_ncon=float(np.mean(ncon)),
                   states_in_contact=int(np.sum(np.array(ncon) > 0)), cases=cases)
        with open(os.path.join(out_dir, f"dense_{name}.json"), "w") as f:
            json.dump(doc, f)
        print(f"{name}: {n} states, {doc['states_in_contact']} in contact, mean ncon {doc['mean_ncon']:.2f}")


if __name__ == "__main__":
    main()
