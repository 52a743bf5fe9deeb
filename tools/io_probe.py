"""Where does ox_batch_step_io spend its time? (pinned host buffers; per-call wall clock, 300 calls)"""
import sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oxide_control_b200 as ox
from oxide_control_b200 import _abi as A
from support import random_state
name, nenv = sys.argv[1], int(sys.argv[2])
m = ox.Model.from_xml_string(getattr(ox.models, name.upper()))
qpos, qvel = random_state(m, nenv, seed=1)
b = ox.BatchedPhysics(m, nenv, precision="f32")
b.set("qpos", qpos); b.set("qvel", qvel)
f32 = torch.float32
act = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (nenv, m.nu)).astype(np.float32)).pin_memory()
q = torch.empty(nenv, m.nq, dtype=f32).pin_memory(); v = torch.empty(nenv, m.nv, dtype=f32).pin_memory()
def run(label, fn, n=300):
    for _ in range(30): fn()
    b.sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    b.sync(); dt = (time.perf_counter() - t0) / n
    print(f"{label:50s} {dt*1e3:.4f} ms/step  {nenv/dt/1e6:.2f} M/s", flush=True)
run("step(1) + sync", lambda: (b.step(1), b.sync()))
run("step_io(ctrl, -, -)", lambda: (b.step_io_ptr(act.data_ptr(), None, None, A.F32, A.MEM_HOST), b.sync()))
run("step_io(-, qpos, qvel)", lambda: b.step_io_ptr(None, q.data_ptr(), v.data_ptr(), A.F32, A.MEM_HOST))
run("step_io(ctrl, qpos, qvel)", lambda: b.step_io_ptr(act.data_ptr(), q.data_ptr(), v.data_ptr(), A.F32, A.MEM_HOST))
run("set(ctrl) + step + get_many(qpos,qvel)", lambda: (b.set_ptr("ctrl", act.data_ptr(), A.F32, A.MEM_HOST, A.LAYOUT_ENV_MAJOR), b.step(1),
                                                       b.get_many_ptr(("qpos", "qvel"), (q.data_ptr(), v.data_ptr()), A.F32, A.MEM_HOST, A.LAYOUT_ENV_MAJOR)))
