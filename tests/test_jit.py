"""Run-time specialisation (csrc/ox_jit.cpp): the model-specialised step kernel for models that were NOT compiled into the
library - Physics::from_xml_string accepts any model at run time (/root/reference/src/physics.rs:18-24), so the fast path
must too. CPU: the unit is generated and compiled to a cubin with nvcc (no GPU needed) and cached. GPU: the cubin is
loaded, reports a specialised kernel name, and passes the same fp64 parity gates against the oracle."""
import os
import time

import numpy as np
import pytest

import oxide_control_b200 as ox
from support import OracleData, SEED, random_state, rel_err
from zoo_models import HOPPER

CHEETAH_DT8 = ox.models.CHEETAH.replace('<option timestep="0.01"/>', '<option timestep="0.008"/>')
USER_MODELS = {"cheetah_dt8": CHEETAH_DT8, "hopper": HOPPER}


def test_variant_really_differs_from_the_compiled_in_model():
    assert CHEETAH_DT8 != ox.models.CHEETAH
    assert ox.Model.from_xml_string(CHEETAH_DT8).timestep == 0.008


@pytest.mark.parametrize("name", list(USER_MODELS))
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_jit_compiles_a_user_model_without_a_gpu(name, precision):
    m = ox.Model.from_xml_string(USER_MODELS[name])
    path = m.jit_compile(precision)          # nvcc -cubin for sm_100a; cached under $OX_B200_CACHE_DIR
    assert os.path.exists(path) and path.endswith(".cubin") and precision in os.path.basename(path)
    blob = open(path, "rb").read()
    assert len(blob) > 50_000 and b"ox_jit_step" in blob
    t0 = time.time()
    assert m.jit_compile(precision) == path  # second request: cache hit, no compiler run
    assert time.time() - t0 < 2.0


def test_jit_can_be_disabled(monkeypatch):
    monkeypatch.setenv("OX_B200_JIT", "0")
    m = ox.Model.from_xml_string(HOPPER)
    with pytest.raises(ox.Error, match="disabled"):
        m.jit_compile("f32")


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(USER_MODELS))
def test_jit_kernel_parity_fp64(name):
    m = ox.Model.from_xml_string(USER_MODELS[name])
    nenv, nsteps = 128, 30
    qpos, qvel = random_state(m, nenv, seed=31)
    if name == "hopper":
        qpos[:, 1] -= 0.12          # foot into the floor: contacts from the first step
    b = ox.BatchedPhysics(m, nenv, precision="f64", specialize=2)
    assert b.kernel_name().startswith("jit_"), (b.kernel_name(), b.jit_note())
    g = ox.BatchedPhysics(m, nenv, precision="f64", specialize=0)
    assert "generic" in g.kernel_name()
    for x in (b, g):
        x.set("qpos", qpos); x.set("qvel", qvel); x.ctrl_philox(True, SEED)
    b.step(1); b.sync()
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        od.fill_ctrl_philox(e, 0); od.step()
        ods.append(od)
    for f in ("qpos", "qvel", "qacc", "sensordata", "ctrl"):
        assert rel_err(b.get(f), np.stack([od.field(f) for od in ods])) <= 1e-9, f
    assert np.array_equal(b.get("ncon")[:, 0], [od.int("ncon") for od in ods])
    for s in range(1, nsteps):
        for e, od in enumerate(ods):
            od.fill_ctrl_philox(e, s); od.step()
    b.step(nsteps - 1); b.sync()
    assert sum(od.int("ncon") for od in ods) > 0
    assert rel_err(b.get("qpos"), np.stack([od.field("qpos") for od in ods])) <= 1e-6
    assert int(b.diverged().sum()) == 0
    # one launch per step, like the compiled-in kernels
    n0 = b.launch_count(); b.step(1); b.step(1); b.sync()
    assert b.launch_count() - n0 == 2


@pytest.mark.gpu
def test_default_config_specialises_large_batches_of_unknown_models():
    m = ox.Model.from_xml_string(CHEETAH_DT8)
    small = ox.BatchedPhysics(m, 64, precision="f32")            # specialize = 1: small batches keep the generic kernel
    assert "generic" in small.kernel_name()
    big = ox.BatchedPhysics(m, 2048, precision="f32")            # >= 1024 envs: specialised at run time (cubin from the cache)
    assert big.kernel_name().startswith("jit_"), big.jit_note()
    big.ctrl_philox(True, SEED); big.step(50); big.sync()
    assert np.isfinite(big.get("qpos")).all() and int(big.diverged().sum()) == 0


@pytest.mark.gpu
def test_jit_kernel_speed_matches_the_compiled_in_one():
    """cheetah with another timestep runs a run-time compiled kernel: same code generator, so within 25 % of the built-in one."""
    import torch
    times = {}
    for name, xml in (("builtin", ox.models.CHEETAH), ("jit", CHEETAH_DT8)):
        m = ox.Model.from_xml_string(xml)
        b = ox.BatchedPhysics(m, 8192, precision="f32", specialize=2)
        qpos, qvel = random_state(m, 8192, seed=32)
        b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
        b.step(100); b.sync()
        t0 = time.perf_counter()
        for _ in range(200):
            b.step(1)
        b.sync()
        times[name] = (time.perf_counter() - t0) / 200
        assert b.kernel_name().startswith("cheetah" if name == "builtin" else "jit_")
    print("ms/step", {k: 1e3 * v for k, v in times.items()})
    assert times["jit"] <= 1.25 * times["builtin"]
