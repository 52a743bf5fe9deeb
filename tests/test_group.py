"""ox_group: one batch per GPU driven from one host process, envs sharded by global env id, the statistics
all-reduce as the only exchange. A single-GPU box exercises the whole code path with a group of one (and of two batches on
the same device); tools/group_bench.py runs it across the GPUs of a multi-GPU box."""
import ctypes as C

import numpy as np
import pytest

import oxide_control_b200 as ox
from oxide_control_b200 import _abi as A
from support import SEED, random_state


def test_group_create_refuses_without_device_or_with_bad_arguments():
    L = A.lib()
    m = ox.Model.from_xml_string(ox.models.CARTPOLE)
    cfg = A.BatchConfig(); L.ox_batch_config_default(C.byref(cfg)); cfg.nenv = 16
    g = C.c_void_p()
    assert L.ox_group_create(None, C.byref(cfg), 1, None, C.byref(g)) == A.OX_ERR_INVALID
    assert L.ox_group_create(m.handle, C.byref(cfg), 0, None, C.byref(g)) == A.OX_ERR_INVALID
    import torch
    if not torch.cuda.is_available():
        assert L.ox_group_create(m.handle, C.byref(cfg), 1, None, C.byref(g)) == A.OX_ERR_CUDA    # no CPU fallback
        assert b"no CUDA device" in L.ox_last_error_message()
    assert L.ox_group_step(None, 1) == A.OX_ERR_INVALID and L.ox_group_size(None) == -1


@pytest.mark.gpu
def test_group_shards_by_global_env_id_and_reduces_stats():
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    per, nsteps = 256, 60
    qpos, qvel = random_state(m, 2 * per, seed=61)
    # two "ranks" (both on device 0 here; on a multi-GPU box devices=[0,1]) against one batch holding all envs
    g = ox.PhysicsGroup(m, per, 2, precision="f32", devices=[0, 0])
    assert g.size == 2 and g.stats_backend() in ("nccl", "host")
    full = ox.BatchedPhysics(m, 2 * per, precision="f32")
    full.set("qpos", qpos); full.set("qvel", qvel); full.ctrl_philox(True, SEED)
    for r, b in enumerate(g.batches):
        b.set("qpos", qpos[r * per:(r + 1) * per]); b.set("qvel", qvel[r * per:(r + 1) * per])
    g.ctrl_philox(True, SEED)
    g.step(nsteps); g.sync()
    full.step(nsteps); full.sync()
    got = np.concatenate([b.get("qpos") for b in g.batches])
    assert np.array_equal(got, full.get("qpos"))            # same kernel, same global env ids -> bit-identical trajectories
    st, ref = g.stats(), full.stats()
    assert st["sum_ncon"] == ref["sum_ncon"] > 0 and st["sum_nefc"] == ref["sum_nefc"] and st["sum_niter"] == ref["sum_niter"]
    assert g.stats()["sum_ncon"] == 0                        # accumulators were cleared by the first call
    g.reset(); g.sync()
    assert not g.batches[1].get("qvel").any()
    g.close()


@pytest.mark.gpu
def test_philox_scale_and_fma_peak():
    m = ox.Model.from_xml_string(ox.models.HUMANOID)
    b = ox.BatchedPhysics(m, 64, precision="f32")
    b.ctrl_philox(True, SEED); b.ctrl_philox_scale(0.125); b.step(1); b.sync()
    c = b.get("ctrl")
    b2 = ox.BatchedPhysics(m, 64, precision="f32")
    b2.ctrl_philox(True, SEED); b2.step(1); b2.sync()
    assert np.array_equal(c, 0.125 * b2.get("ctrl")) and np.abs(c).max() <= 0.125
    from support import OracleData, oracle_lib
    oracle_lib().oxo_set_ctrl_scale(C.c_double(0.125))
    try:
        od = OracleData(m); od.fill_ctrl_philox(5, 0)
        assert np.array_equal(od.field("ctrl"), c[5].astype(np.float64))     # CPU fp64 and GPU fp32 controls stay bit-identical
    finally:
        oracle_lib().oxo_set_ctrl_scale(C.c_double(1.0))
    p32, p64 = ox.measure_fma_peak(0, "f32"), ox.measure_fma_peak(0, "f64")
    print(f"measured FMA peak: fp32 {p32:.1f} TFLOP/s, fp64 {p64:.1f} TFLOP/s")
    assert 30 < p32 < 90 and 15 < p64 < 50                                   # B200: 148 SMs x 128 (64) lanes x 2 x ~1.9 GHz
