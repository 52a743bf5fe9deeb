#!/usr/bin/env python
"""Golden-vector hook. Run this OUTSIDE the build environment, where `pip install mujoco==3.3.2`
works, from the repo root:

    python tools/dump_mujoco_golden.py            # writes tests/golden/<config>.json

For every synthetic model of oxide_control_b200/models.py and a fixed seed it records, from the real MuJoCo C engine
(the arithmetic behind oxide_control's Physics::step, reference src/physics.rs:44-46), the inputs
(qpos, qvel, ctrl) and the outputs after 1 and 100 steps (qpos, qvel, qacc, sensordata) plus stage outputs of the first
forward (qM, qfrc_bias, ncon, nefc, efc_D, efc_aref, efc_J). tests/test_golden_hook.py compares the oracle (and, on a
GPU, the CUDA path) with these files when they exist and reports SKIPPED when they do not - it never passes silently.
"""
import importlib.util
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_models():
    spec = importlib.util.spec_from_file_location("ox_models", os.path.join(ROOT, "oxide_control_b200", "models.py"))  # reads spec_models/*.xml
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    import mujoco  # noqa: raises here in the build environment, by design
    assert mujoco.__version__.startswith("3.3."), mujoco.__version__
    models = load_models()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, cfg in models.CONFIGS.items():
        m = mujoco.MjModel.from_xml_string(cfg["xml"])
        d = mujoco.MjData(m)
        rng = np.random.default_rng(20261018)
        cases = []
        for case in range(4):
            mujoco.mj_resetData(m, d)
            d.qpos[:] = m.qpos0 + 0.0
            for j in range(m.njnt):
                a = m.jnt_qposadr[j]
                if m.jnt_type[j] in (2, 3):
                    d.qpos[a] += rng.uniform(-0.1, 0.1)
                elif m.jnt_type[j] == 0:
                    d.qpos[a + 2] += rng.uniform(0, 0.2)
            d.qvel[:] = rng.normal(0, 0.1, m.nv)
            ctrl = rng.uniform(-1, 1, (100, m.nu))
            rec = {"qpos": d.qpos.tolist(), "qvel": d.qvel.tolist(), "ctrl": ctrl.tolist(), "after": {}}
            for s in range(100):
                d.ctrl[:] = ctrl[s]
                mujoco.mj_step(m, d)
                if s == 0:
                    qM = np.zeros((m.nv, m.nv))
                    mujoco.mj_fullM(m, qM, d.qM)
                    rec["first_forward"] = {
                        "M": qM.tolist(), "qfrc_bias": d.qfrc_bias.tolist(), "qacc_smooth": d.qacc_smooth.tolist(),
                        "ncon": int(d.ncon), "nefc": int(d.nefc), "efc_D": d.efc_D.tolist(), "efc_aref": d.efc_aref.tolist(),
                        "efc_J": np.asarray(d.efc_J).reshape(d.nefc, m.nv).tolist() if d.nefc and not mujoco.mj_isSparse(m) else None,
                        "con_dist": [float(c.dist) for c in d.contact[:d.ncon]],
                    }
                if s in (0, 99):
                    rec["after"][str(s + 1)] = {"qpos": d.qpos.tolist(), "qvel": d.qvel.tolist(), "qacc": d.qacc.tolist(),
                                                "sensordata": d.sensordata.tolist()}
            cases.append(rec)
        consts = {"body_mass": m.body_mass.tolist(), "body_inertia": m.body_inertia.tolist(), "dof_invweight0": m.dof_invweight0.tolist(),
                  "body_invweight0": m.body_invweight0.tolist(), "meaninertia": float(m.stat.meaninertia), "nM": int(m.nM)}
        with open(os.path.join(out_dir, f"{name}.json"), "w") as f:
            json.dump({"mujoco": mujoco.__version__, "model": name, "constants": consts, "cases": cases}, f)
        print("wrote", name)


if __name__ == "__main__":
    sys.exit(main())
