"""world_size-2 gloo run of the host-side multi-GPU logic: env sharding is trajectory-stable and the
only collective is the statistics reduction. The physics inside each rank runs on the test-only host instantiation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oxide_control_b200 as ox
    from oxide_control_b200.sharding import gather_stats, max_over_ranks, shard_range
    from support import HostBatch, SEED
    from bench import initial_state
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    per = 8
    lo, hi = shard_range(rank, world, per)
    q, v = initial_state(m, world * per, lo, hi)
    hb = HostBatch(m, per, "f64")
    hb.set("qpos", q); hb.set("qvel", v)
    nsteps = 30
    hb.step(nsteps, True, SEED, lo, 0)
    np.save(os.path.join(out_dir, f"q{rank}.npy"), hb.get("qpos"))
    tot = gather_stats({"env_steps": per * nsteps, "sum_ncon": float(hb.get("ncon").sum()), "diverged": float(hb.get("diverged").sum())})
    t = max_over_ranks(1.0 + rank)
    if rank == 0:
        np.save(os.path.join(out_dir, "tot.npy"), np.array([tot["env_steps"], tot["sum_ncon"], tot["diverged"], t]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    sys.path.insert(0, ROOT)
    import oxide_control_b200 as ox
    from support import HostBatch, SEED
    from bench import initial_state
    world, per, nsteps = 2, 8, 30
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    q, v = initial_state(m, world * per, 0, world * per)
    hb = HostBatch(m, world * per, "f64")
    hb.set("qpos", q); hb.set("qvel", v)
    hb.step(nsteps, True, SEED, 0, 0)
    full = hb.get("qpos")
    got = np.concatenate([np.load(tmp_path / f"q{r}.npy") for r in range(world)])
    assert np.array_equal(full, got)                       # bit-identical per global env id
    tot = np.load(tmp_path / "tot.npy")
    assert tot[0] == world * per * nsteps and tot[2] == 0 and tot[3] == 2.0
    assert tot[1] == float(hb.get("ncon").sum())


def test_shard_helpers():
    from oxide_control_b200.sharding import owner_of, shard_range
    assert shard_range(3, 8, 8192) == (24576, 32768) and owner_of(24577, 8192) == (3, 1)
