"""fp32 throughput mode: what replaces north_star's flat 1e-4 gate on qacc where it is unattainable, and why.

Measured on B200 (tools/fp32_study.py -> profiles/r2_fp32_study.jsonl), states in contact, rel = |a-b| / max(1,|b|):
  cheetah : forward error of qacc vs the fp64 oracle max 5.1e-4, cond_inf(H) up to 6.3e3, BACKWARD error max 6.9e-6
  humanoid: forward error max 1.7e-3, cond_inf(H) up to 4.6e5, backward error max 6.6e-6
  and two fp32 evaluations of the very same stage templates (GPU vs host instantiation, differing only in FMA contraction)
  differ from EACH OTHER by 8.8e-5 / 1.9e-3 - as much as either differs from fp64.
So the forward gap is cond(H) * eps, a property of the problem, not of the kernel. The gates below therefore are:
  (1) backward stability: the fp32 qacc satisfies the fp64 stationarity condition to a normwise residual <= 3e-5
      (a few tens of fp32 ulps plus the solver tolerance of 1e-6);
  (2) forward error <= 50 * cond(H) * backward error, env by env (the textbook bound with a constant of 50);
  (3) north_star's "bounded divergence over short horizons" in fp32 for the two contact configs, against the fp64 oracle:
      median / 90th percentile over envs of max_t |qpos_fp32 - qpos_fp64| over 100 steps, no auto-resets.
The flat per-config forward gates of tests/test_gpu_parity.py::FP32_TOL stay as regression guards."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from fp32_metrics import backward_error
from support import HostBatch, OracleData, SEED, random_state, rel_err

pytestmark = pytest.mark.gpu

NENV = 96
BACKWARD_MAX = 3e-5
HORIZON = {"cheetah": dict(median=5e-5, p90=5e-4), "humanoid": dict(median=5e-4, p90=5e-3)}


def settled_states(m, nenv, warm=60):
    qpos, qvel = random_state(m, nenv, seed=12)
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(warm):
            od.fill_ctrl_philox(e, s); od.step()
        ods.append(od)
    r32 = lambda f: np.stack([od.field(f) for od in ods]).astype(np.float32).astype(np.float64)
    return ods, r32("qpos"), r32("qvel"), r32("qacc_warmstart")


@pytest.mark.parametrize("name", ["cartpole", "cheetah", "humanoid"])
def test_fp32_step_is_backward_stable_and_forward_error_is_conditioning(name):
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    ods, q, v, w = settled_states(m, NENV)
    b = ox.BatchedPhysics(m, NENV, precision="f32")
    hb = HostBatch(m, NENV, "f32")
    for x in (b, hb):
        x.set("qpos", q); x.set("qvel", v); x.set("qacc_warmstart", w)
    b.ctrl_philox(True, SEED); b.set_step_counter(60); b.step(1); b.sync()
    hb.step(1, True, SEED, 0, 60)
    got = b.get("qacc")
    eta_max, ratio_max, fwd_max, ref = 0.0, 0.0, 0.0, []
    for e, od in enumerate(ods):
        od.field("qpos")[:] = q[e]; od.field("qvel")[:] = v[e]; od.field("qacc_warmstart")[:] = w[e]
        od.fill_ctrl_philox(e, 60); od.forward()
        a64 = od.field("qacc").copy(); ref.append(a64)
        eta, cond = backward_error(m, od, got[e])
        fwd = float(np.max(np.abs(got[e] - a64) / np.maximum(1, np.abs(a64))))
        eta_max, fwd_max = max(eta_max, eta), max(fwd_max, fwd)
        assert fwd <= 50 * cond * max(eta, 6e-8), (e, fwd, cond, eta)
    two_fp32 = rel_err(got, hb.get("qacc"))
    print(f"{name}: fp32 qacc forward err max {fwd_max:.2e}, backward err max {eta_max:.2e}, GPU fp32 vs host fp32 {two_fp32:.2e}")
    assert eta_max <= BACKWARD_MAX
    if name != "cartpole":
        assert sum(od.int("ncon") for od in ods) > 0.3 * NENV      # the states really are in contact


@pytest.mark.parametrize("name", ["cheetah", "humanoid"])
def test_fp32_horizon_bounded_divergence_with_contacts(name):
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    ods, q, v, w = settled_states(m, NENV)
    b = ox.BatchedPhysics(m, NENV, precision="f32")
    b.set("qpos", q); b.set("qvel", v); b.set("qacc_warmstart", w); b.ctrl_philox(True, SEED); b.set_step_counter(60)
    for e, od in enumerate(ods):
        od.field("qpos")[:] = q[e]; od.field("qvel")[:] = v[e]; od.field("qacc_warmstart")[:] = w[e]
    dev = np.zeros(NENV)
    for chunk in range(10):
        b.step(10); b.sync()
        for e, od in enumerate(ods):
            for s in range(10):
                od.fill_ctrl_philox(e, 60 + 10 * chunk + s); od.step()
        dev = np.maximum(dev, np.max(np.abs(b.get("qpos") - np.stack([od.field("qpos") for od in ods])), axis=1))
    print(f"{name}: 100-step fp32 vs fp64 max|dqpos| median {np.median(dev):.2e} p90 {np.percentile(dev, 90):.2e} max {dev.max():.2e}")
    assert np.median(dev) <= HORIZON[name]["median"] and np.percentile(dev, 90) <= HORIZON[name]["p90"]
    assert np.isfinite(b.get("qpos")).all() and int(b.diverged().sum()) == 0
