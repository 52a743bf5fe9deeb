#!/usr/bin/env python
"""Single-process multi-GPU run through ox_group (no torchrun, no torch.distributed): N batches, one per GPU, envs sharded
by global env id, stepped by one host thread per device; statistics summed by ncclAllReduce inside the library.
    python tools/group_bench.py --gpus 2 --steps 500
Prints one JSON line. bench.py (one process per GPU under torchrun) remains the contract benchmark; this shows the same
scaling from a single host process - what a Rust host of the reference would do."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oxide_control_b200 as ox  # noqa: E402
from bench import SEED, initial_state  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--config", default="cheetah")
    args = ap.parse_args()
    cfg = ox.models.CONFIGS[args.config]
    m = ox.Model.from_xml_string(cfg["xml"])
    per = cfg["nenv"]
    g = ox.PhysicsGroup(m, per, args.gpus, precision=cfg["precision"])
    for r, b in enumerate(g.batches):
        q, v = initial_state(m, args.gpus * per, r * per, (r + 1) * per)
        b.set("qpos", q); b.set("qvel", v)
    g.ctrl_philox(True, SEED)
    g.step(args.warmup); g.sync(); g.stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.step(1)
    g.sync()
    dt = time.perf_counter() - t0
    st = g.stats()
    n = args.gpus * per * args.steps
    print(json.dumps({"metric": "env-steps/sec", "value": n / dt, "n_gpus": args.gpus, "steps": args.steps, "ms_per_step": 1e3 * dt / args.steps,
                      "driver": "ox_group (one process, one host thread per GPU)", "stats_backend": g.stats_backend(),
                      "mean_ncon": st["sum_ncon"] / n, "mean_solver_iters": st["sum_niter"] / n, "diverged": st["diverged"],
                      "timing": "host wall clock around K x ox_group_step(1) + ox_group_sync"}))


if __name__ == "__main__":
    main()
