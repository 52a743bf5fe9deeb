#!/usr/bin/env python
"""Small driver for profiling: L launches of S steps each on one config (used under ncu and for launch-shape experiments)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oxide_control_b200 as ox  # noqa: E402
from bench import initial_state  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cheetah")
ap.add_argument("--nenv", type=int, default=8192)
ap.add_argument("--precision", default="f32")
ap.add_argument("--mode", default="fused")
ap.add_argument("--steps-per-launch", type=int, default=1)
ap.add_argument("--launches", type=int, default=20)
ap.add_argument("--warmup", type=int, default=100)
ap.add_argument("--iterations", type=int, default=0)
ap.add_argument("--ls-iterations", type=int, default=0)
ap.add_argument("--block", type=int, default=0)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--no-spec", action="store_true")
ap.add_argument("--lanes", type=int, default=0)
ap.add_argument("--coop", type=int, default=-1)
a = ap.parse_args()

import torch
from oxide_control_b200 import _abi as A
model = ox.Model.from_xml_string(ox.models.CONFIGS[a.config]["xml"])
b = ox.BatchedPhysics(model, a.nenv, precision=a.precision, mode=a.mode, iterations=a.iterations, ls_iterations=a.ls_iterations,
                      block_threads=a.block, use_graph=a.graph, specialize=not a.no_spec, lanes_per_warp=a.lanes, coop_solver=a.coop)
q, v = initial_state(model, a.nenv, 0, a.nenv)
b.set("qpos", q); b.set("qvel", v); b.ctrl_philox(True, 0x0B200)
b.step(a.warmup); b.sync(); b.stats()
stream = torch.cuda.ExternalStream(A.lib().ox_batch_stream(b.handle))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
with torch.cuda.stream(stream):
    e0.record(stream)
    for _ in range(a.launches):
        b.step(a.steps_per_launch)
    e1.record(stream)
b.sync()
ms = e0.elapsed_time(e1)
n = a.launches * a.steps_per_launch
st = b.stats()
den = a.nenv * n
print(f"[{b.kernel_name()}] {a.config} {a.precision} nenv={a.nenv} mode={a.mode} steps/launch={a.steps_per_launch} launches={a.launches} it={a.iterations} ls={a.ls_iterations} block={a.block} lanes={a.lanes}: "
      f"{ms / n:.4f} ms/step  {a.nenv * n / ms * 1e3:.3e} env-steps/s  ncon={st['sum_ncon'] / den:.2f} nefc={st['sum_nefc'] / den:.2f} "
      f"iters={st['sum_niter'] / den:.2f} div={st['diverged']}", flush=True)
