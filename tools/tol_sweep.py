"""fp32 solver tolerance sweep: step time, mean solver iterations and single-step error vs the fp64 oracle (cheetah)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oxide_control_b200 as ox
from oxide_control_b200 import _abi as A
from support import OracleData, SEED, random_state, rel_err

cfgname = sys.argv[1] if len(sys.argv) > 1 else "cheetah"
nenv = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
m = ox.Model.from_xml_string(getattr(ox.models, cfgname.upper()))
qpos, qvel = random_state(m, nenv, seed=1)
for tol in (-1.0, 1e-7, 1e-6, 3e-6, 1e-5):
    b = ox.BatchedPhysics(m, nenv, precision="f32", tolerance=tol)
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    b.step(200); b.sync(); b.stats()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.ExternalStream(A.lib().ox_batch_stream(b.handle))
    s.record(stream)
    for _ in range(300): b.step(1)
    e.record(stream); b.sync()
    ms = s.elapsed_time(e) / 300
    st = b.stats()
    # single-step error vs oracle from the current state, 128 envs
    q, v, w = b.get("qpos")[:128], b.get("qvel")[:128], b.get("qacc_warmstart")[:128]
    t = b.get("time")[:128]
    b.step(1); b.sync()
    c = b.get("ctrl")[:128]
    errs = []
    for i in range(128):
        od = OracleData(m)
        od.field("qpos")[:] = q[i]; od.field("qvel")[:] = v[i]; od.field("qacc_warmstart")[:] = w[i]; od.field("ctrl")[:] = c[i]; od.field("time")[:] = t[i]
        od.step()
        errs.append((rel_err(b.get1("qvel", i), od.field("qvel")), rel_err(b.get1("qacc", i), od.field("qacc"))))
    errs = np.array(errs)
    print(f"tol {tol:8.1e}  {ms:.4f} ms/step  {nenv/ms/1e3:.1f} M/s  iters {st['sum_niter']/(300*nenv):.3f}  qvel err max {errs[:,0].max():.2e}  qacc err max {errs[:,1].max():.2e} mean {errs[:,1].mean():.2e}", flush=True)
    b.close()
