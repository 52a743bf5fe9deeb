#!/usr/bin/env python
"""Small run of every kernel family for compute-sanitizer (memcheck / initcheck): model-specialised fused, generic fused,
staged (thread-per-env solver), staged + warp-cooperative solver, split pipeline (humanoid), run-time specialised (jit),
env layer, state snapshot, group statistics - on the small configs, a few steps each. Exits non-zero on any API error."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("OX_B200_CACHE_DIR", os.path.join(ROOT, "build", "jit_cache"))
import oxide_control_b200 as ox  # noqa: E402
from support import SEED, random_state  # noqa: E402
from zoo_models import HOPPER, ZOO  # noqa: E402

CASES = [
    ("cheetah", ox.models.CHEETAH, dict(mode="fused"), "f32"), ("cheetah", ox.models.CHEETAH, dict(mode="fused"), "f64"),
    ("cheetah", ox.models.CHEETAH, dict(mode="fused", specialize=0), "f32"), ("cheetah", ox.models.CHEETAH, dict(mode="staged", coop_solver=0), "f32"),
    ("cheetah", ox.models.CHEETAH, dict(mode="staged", coop_solver=1), "f32"), ("humanoid", ox.models.HUMANOID, dict(mode="fused"), "f32"),
    ("humanoid", ox.models.HUMANOID, dict(mode="staged"), "f64"), ("acrobot", ox.models.ACROBOT, dict(mode="fused"), "f64"),
    ("zoo_a", ZOO["zoo_a"], dict(mode="fused"), "f64"), ("zoo_b", ZOO["zoo_b"], dict(mode="staged"), "f32"), ("zoo_c", ZOO["zoo_c"], dict(mode="fused"), "f32"),
    ("hopper", HOPPER, dict(mode="fused", specialize=2), "f32"),
]
for name, xml, kw, prec in CASES:
    m = ox.Model.from_xml_string(xml)
    nenv = 96
    b = ox.BatchedPhysics(m, nenv, precision=prec, **kw)
    q, v = random_state(m, nenv, seed=7)
    b.set("qpos", q); b.set("qvel", v); b.ctrl_philox(True, SEED)
    b.step(3); b.step(1); b.forward(); b.sync()
    st = b.get_state(); b.set_state(st); b.step(2); b.sync()
    assert np.isfinite(b.get("qpos")).all()
    print(f"ok {name} {prec} {kw} kernel={b.kernel_name()}", flush=True)
m = ox.Model.from_xml_string(ox.models.CHEETAH)
b = ox.BatchedPhysics(m, 64, precision="f32")
env = ox.BatchedEnvironment(b, ox.TaskSpec(obs=[("qpos", 0, m.nq), ("qvel", 0, m.nv)], reward=[("qvel", 0, "linear", 1.0)], time_limit=0.05,
                                           init_qpos_noise=0.1, init_qvel_noise=0.1, seed=3))
env.reset()
for _ in range(8):
    env.step(np.zeros((64, m.nu), dtype=np.float32))
env.close()
g = ox.PhysicsGroup(m, 64, 2, devices=[0, 0])
g.ctrl_philox(True, SEED); g.step(4); g.sync(); print("group stats", g.stats(), g.stats_backend()); g.close()
print("sanitize_run done")
