#!/usr/bin/env python
"""fp32 accuracy study (GPU): forward error of the fp32 step against the fp64 oracle, its normwise BACKWARD error in the fp64
equations, the conditioning that links the two, the host fp32 instantiation of the same templates, and 100-step divergence.
Prints one JSON line per config; tests/test_gpu_fp32.py asserts the bounds derived from it."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oxide_control_b200 as ox  # noqa: E402
from support import HostBatch, OracleData, SEED, random_state, rel_err  # noqa: E402
from fp32_metrics import backward_error, dense_M  # noqa: E402


def main():
    for name in ("cartpole", "cheetah", "humanoid"):
        m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
        nenv = 128
        qpos, qvel = random_state(m, nenv, seed=12)
        # settle into contact on the fp64 oracle first, so that the single-step comparison is on constrained states
        ods = []
        for e in range(nenv):
            od = OracleData(m)
            od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
            for s in range(60):
                od.fill_ctrl_philox(e, s); od.step()
            ods.append(od)
        q32 = np.stack([od.field("qpos") for od in ods]).astype(np.float32).astype(np.float64)
        v32 = np.stack([od.field("qvel") for od in ods]).astype(np.float32).astype(np.float64)
        w32 = np.stack([od.field("qacc_warmstart") for od in ods]).astype(np.float32).astype(np.float64)
        b = ox.BatchedPhysics(m, nenv, precision="f32")
        hb = HostBatch(m, nenv, "f32")
        for x in (b, hb):
            x.set("qpos", q32); x.set("qvel", v32); x.set("qacc_warmstart", w32)
        b.ctrl_philox(True, SEED); b.set_step_counter(60); b.step(1); b.sync()
        hb.step(1, True, SEED, 0, 60)
        fwd, eta, cond, ncon = [], [], [], 0
        ref_a = []
        for e, od in enumerate(ods):
            od.field("qpos")[:] = q32[e]; od.field("qvel")[:] = v32[e]; od.field("qacc_warmstart")[:] = w32[e]
            od.fill_ctrl_philox(e, 60); od.forward()
            a64 = od.field("qacc").copy(); ref_a.append(a64)
            a32 = b.get("qacc")[e]
            be, kappa = backward_error(m, od, a32)
            eta.append(be); cond.append(kappa); ncon += od.int("ncon")
            fwd.append(np.max(np.abs(a32 - a64) / np.maximum(1, np.abs(a64))))
        ref_a = np.array(ref_a)
        out = dict(config=name, mean_ncon=ncon / nenv, fwd_err_max=float(np.max(fwd)), fwd_err_median=float(np.median(fwd)),
                   backward_err_max=float(np.max(eta)), backward_err_median=float(np.median(eta)), cond_max=float(np.max(cond)),
                   cond_median=float(np.median(cond)), ratio_fwd_over_cond_eta_max=float(np.max(np.array(fwd) / (np.array(cond) * np.array(eta) + 1e-30))),
                   gpu_vs_host_f32_qacc=rel_err(b.get("qacc"), hb.get("qacc")), host_f32_vs_f64_qacc=rel_err(hb.get("qacc"), ref_a))
        # 100-step horizon, fp32 GPU vs fp64 oracle from the same (fp32-representable) state
        b2 = ox.BatchedPhysics(m, nenv, precision="f32")
        b2.set("qpos", q32); b2.set("qvel", v32); b2.set("qacc_warmstart", w32); b2.ctrl_philox(True, SEED); b2.set_step_counter(60)
        dev = np.zeros((nenv,))
        for chunk in range(10):
            b2.step(10); b2.sync()
            for e, od in enumerate(ods):
                if chunk == 0:
                    od.field("qpos")[:] = q32[e]; od.field("qvel")[:] = v32[e]; od.field("qacc_warmstart")[:] = w32[e]
                for s in range(10):
                    od.fill_ctrl_philox(e, 60 + 10 * chunk + s); od.step()
            ref = np.stack([od.field("qpos") for od in ods])
            dev = np.maximum(dev, np.max(np.abs(b2.get("qpos") - ref), axis=1))
        out.update(horizon100_qpos_dev_median=float(np.median(dev)), horizon100_qpos_dev_p90=float(np.percentile(dev, 90)),
                   horizon100_qpos_dev_max=float(np.max(dev)), diverged=int(b2.diverged().sum()))
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
