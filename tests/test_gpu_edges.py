"""GPU edge cases and size-independent properties at BASELINE.json's full batch sizes: ragged and tiny
batches, multi-step launches vs repeated single steps, determinism, env independence at 8192 / 4096 envs, masked reset,
CUDA-graph replay of the staged sequence, argument errors. All through the C ABI; the oracle checks the small cases."""
import numpy as np
import pytest

from support import OracleData, SEED, random_state, rel_err

pytestmark = pytest.mark.gpu


def _oracle_final(model, qpos, qvel, nsteps, env_off=0):
    out = []
    for e in range(qpos.shape[0]):
        od = OracleData(model)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(env_off + e, s)
            od.step()
        out.append(np.concatenate([od.field("qpos"), od.field("qvel"), od.field("qacc")]))
    return np.stack(out)


@pytest.mark.parametrize("nenv", [1, 31, 33, 100])
@pytest.mark.parametrize("name", ["cheetah", "humanoid"])
def test_ragged_batch_sizes_fp64(ox, name, nenv):
    """Batches that do not fill a warp / a CTA (the arena stride is nenv rounded up to 32): every env is stepped, none is
    corrupted by its padding neighbours."""
    model = ox.Model.from_xml_string(getattr(ox.models, name.upper()))
    qpos, qvel = random_state(model, nenv, seed=21)
    b = ox.BatchedPhysics(model, nenv, precision="f64")
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    b.step(3); b.sync()
    got = np.concatenate([b.get("qpos"), b.get("qvel"), b.get("qacc")], 1)
    assert rel_err(got, _oracle_final(model, qpos, qvel, 3)) <= 1e-9


@pytest.mark.parametrize("name,precision,kw", [("cheetah", "f32", {}), ("humanoid", "f32", {}), ("acrobot", "f64", {}),
                                                ("cheetah", "f32", dict(specialize=False)), ("cheetah", "f64", dict(mode="staged"))])
def test_multi_step_launch_equals_repeated_single_steps_bitwise(ox, name, precision, kw):
    """ox_batch_step(K) (K steps inside one launch, or K graph replays) == K x ox_batch_step(1), bit for bit, including the
    device-side step counter that keys the Philox control stream."""
    model = ox.Model.from_xml_string(getattr(ox.models, name.upper()))
    nenv = 200
    qpos, qvel = random_state(model, nenv, seed=22)
    res = []
    for chunks in ([9], [1] * 9, [4, 5]):
        b = ox.BatchedPhysics(model, nenv, precision=precision, **kw)
        b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
        for k in chunks:
            b.step(k)
        b.sync()
        res.append((b.get("qpos"), b.get("qvel"), b.get("qacc"), b.get("ctrl"), b.get("time")))
    for other in res[1:]:
        for x, y in zip(res[0], other):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("name,nenv", [("cheetah", 8192), ("humanoid", 4096)])
def test_full_size_determinism_and_env_independence(ox, name, nenv):
    """At BASELINE.json's batch sizes: two runs agree bit for bit; envs [nenv-64, nenv) equal a 64-env batch with the same
    global ids (no cross-env coupling anywhere in the step); everything stays finite; contacts really occur; and the first
    8 envs still match the oracle after 50 steps of fp32 contact dynamics to the horizon bound of the parity tests."""
    model = ox.Model.from_xml_string(getattr(ox.models, name.upper()))
    qpos, qvel = random_state(model, nenv, seed=23)
    nsteps = 50
    runs = []
    for _ in range(2):
        b = ox.BatchedPhysics(model, nenv, precision="f32")
        b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
        b.step(nsteps); b.sync()
        runs.append((b.get("qpos"), b.get("qvel"), b.get("qacc")))
        st = b.stats()
    for x, y in zip(*runs):
        assert np.array_equal(x, y)
    assert all(np.isfinite(x).all() for x in runs[0])
    assert st["sum_ncon"] > 0 and st["diverged"] == 0
    tail = ox.BatchedPhysics(model, 64, precision="f32", env_id_offset=nenv - 64)
    tail.set("qpos", qpos[-64:]); tail.set("qvel", qvel[-64:]); tail.ctrl_philox(True, SEED)
    tail.step(nsteps); tail.sync()
    assert np.array_equal(tail.get("qpos"), runs[0][0][-64:]) and np.array_equal(tail.get("qacc"), runs[0][2][-64:])
    want = _oracle_final(model, qpos[:8], qvel[:8], nsteps)
    nq = model.nq
    assert np.abs(runs[0][0][:8] - want[:, :nq]).max() <= 5e-2


def test_masked_reset_touches_only_masked_envs(ox):
    model = ox.Model.from_xml_string(ox.models.HUMANOID)
    nenv = 96
    qpos, qvel = random_state(model, nenv, seed=24)
    b = ox.BatchedPhysics(model, nenv, precision="f64")
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    b.step(5); b.sync()
    before = {f: b.get(f) for f in ("qpos", "qvel", "time", "qacc_warmstart", "ctrl")}
    mask = np.zeros(nenv, np.uint8); mask[[0, 31, 32, 95]] = 1
    b.reset(mask)
    after = {f: b.get(f) for f in before}
    keep = mask == 0
    for f in before:
        assert np.array_equal(after[f][keep], before[f][keep]), f
    assert np.array_equal(after["qpos"][~keep], np.tile(np.array(model.qpos0), (4, 1)))
    for f in ("qvel", "time", "qacc_warmstart", "ctrl"):
        assert not after[f][~keep].any(), f


def test_cuda_graph_replay_of_the_staged_sequence_matches_plain_launches(ox):
    model = ox.Model.from_xml_string(ox.models.CHEETAH)
    nenv = 128
    qpos, qvel = random_state(model, nenv, seed=25)
    out = []
    for graph in (False, True):
        b = ox.BatchedPhysics(model, nenv, precision="f64", mode="staged", use_graph=graph)
        b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
        b.step(6); b.sync()
        out.append((b.get("qpos"), b.get("qacc")))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


def test_argument_errors_are_loud(ox):
    model = ox.Model.from_xml_string(ox.models.PENDULUM)
    b = ox.BatchedPhysics(model, 4, precision="f64")
    q = b.get("qpos")
    b.step(0); b.sync()
    assert np.array_equal(b.get("qpos"), q)                   # zero steps: nothing happens
    with pytest.raises(ox.Error):
        b.step(-1)
    with pytest.raises(ox.Error):
        b.get1("qpos", 4)                                      # env out of range
    with pytest.raises(ox.Error):
        b.get1("qpos", 0, 1, 1)                                # offset out of range
    with pytest.raises(ox.Error):
        ox.BatchedPhysics(model, 0)
    with pytest.raises(ox.Error):
        ox.BatchedPhysics(model, 4, device=99)
    # maximum size: device code indexes every field with 32 bits; a batch whose largest field would overflow is refused
    # before anything is allocated
    big = ox.Model.from_xml_string(ox.models.HUMANOID)
    with pytest.raises(ox.Error, match="32-bit"):
        ox.BatchedPhysics(big, 1 << 22)


def test_pinned_host_buffers_take_the_zero_copy_path_and_agree_with_pageable_ones(ox):
    """Pinned (mapped) host buffers are read / written by the kernels directly over PCIe; pageable ones are staged through
    a device buffer. Same results either way, for bulk set / get / get_many and for the env layer."""
    import torch
    from oxide_control_b200 import _abi as A
    model = ox.Model.from_xml_string(ox.models.CHEETAH)
    nenv = 1000   # ragged: the last warp is partial
    qpos, qvel = random_state(model, nenv, seed=26)
    f32 = torch.float32
    for layout in (A.LAYOUT_ENV_MAJOR, A.LAYOUT_ELEM_MAJOR):
        b = ox.BatchedPhysics(model, nenv, precision="f32")
        shape = (nenv, model.nq) if layout == A.LAYOUT_ENV_MAJOR else (model.nq, nenv)
        src = torch.from_numpy((qpos if layout == A.LAYOUT_ENV_MAJOR else qpos.T).astype(np.float32).copy()).pin_memory()
        b.set_ptr("qpos", src.data_ptr(), A.F32, A.MEM_HOST, layout)
        b.sync()
        assert np.array_equal(b.get("qpos"), qpos.astype(np.float32))            # pageable read-back of a pinned write
        b.set("qvel", qvel); b.ctrl_philox(True, SEED); b.step(3)
        out_q = torch.empty(shape, dtype=f32).pin_memory()
        out_v = torch.empty((nenv, model.nv) if layout == A.LAYOUT_ENV_MAJOR else (model.nv, nenv), dtype=f32).pin_memory()
        b.get_many_ptr(("qpos", "qvel"), (out_q.data_ptr(), out_v.data_ptr()), A.F32, A.MEM_HOST, layout)
        q = out_q.numpy() if layout == A.LAYOUT_ENV_MAJOR else out_q.numpy().T
        v = out_v.numpy() if layout == A.LAYOUT_ENV_MAJOR else out_v.numpy().T
        assert np.array_equal(q, b.get("qpos")) and np.array_equal(v, b.get("qvel"))
    # env layer: pinned vs pageable outputs of the same step
    outs = []
    for pinned in (False, True):
        b = ox.BatchedPhysics(model, nenv, precision="f32")
        env = ox.BatchedEnvironment(b, ox.TaskSpec(obs=[("qpos", 0, 9), ("qvel", 0, 9)], reward=[("qvel", 0, "linear", 1.0)],
                                                   time_limit=0.03, init_qpos_noise=0.1, seed=3))
        env.reset()
        act = np.random.default_rng(0).uniform(-1, 1, (nenv, model.nu)).astype(np.float32)
        if pinned:
            a = torch.from_numpy(act).pin_memory()
            o = torch.empty(nenv, 18, dtype=f32).pin_memory(); r = torch.empty(nenv, dtype=f32).pin_memory()
            d = torch.empty(nenv, dtype=f32).pin_memory(); f = torch.empty(nenv, dtype=torch.uint8).pin_memory()
            for _ in range(4):
                env.step_ptr(a.data_ptr(), o.data_ptr(), r.data_ptr(), d.data_ptr(), f.data_ptr(), A.F32, A.MEM_HOST)
            outs.append((o.numpy().copy(), r.numpy().copy(), d.numpy().copy(), f.numpy().astype(bool)))
        else:
            for _ in range(4):
                ts = env.step(act)
            outs.append((ts.observation, ts.reward, ts.discount, ts.finished))
    for x, y in zip(*outs):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("name,precision,kw", [("cheetah", "f32", {}), ("cheetah", "f64", {}), ("humanoid", "f32", {}), ("cartpole", "f32", {}),
                                                ("cheetah", "f32", dict(specialize=False)), ("cheetah", "f64", dict(mode="staged"))])
def test_step_io_equals_set_step_get(ox, name, precision, kw):
    """ox_batch_step_io (controls in, one step, qpos / qvel out in ONE call; the specialised kernels do the I/O themselves,
    fused or split pipeline, pinned host / device / pageable buffers) == ox_batch_set(ctrl) + ox_batch_step(1) + ox_batch_get."""
    import torch
    from oxide_control_b200 import _abi as A
    model = ox.Model.from_xml_string(getattr(ox.models, name.upper()))
    nenv = 1000  # the last warp is partial
    qpos, qvel = random_state(model, nenv, seed=27)
    rng = np.random.default_rng(5)
    acts = [rng.uniform(-1, 1, (nenv, model.nu)) for _ in range(4)]
    dt = torch.float64 if precision == "f64" else torch.float32
    code = A.F64 if precision == "f64" else A.F32
    ref = ox.BatchedPhysics(model, nenv, precision=precision, **kw)
    ref.set("qpos", qpos); ref.set("qvel", qvel)
    want = []
    for a in acts:
        ref.set("ctrl", a); ref.step(1); ref.sync()
        want.append((ref.get("qpos"), ref.get("qvel")))
    for flavour in ("pageable", "pinned", "device"):
        b = ox.BatchedPhysics(model, nenv, precision=precision, **kw)
        b.set("qpos", qpos); b.set("qvel", qvel)
        for a, (wq, wv) in zip(acts, want):
            if flavour == "pageable":
                q, v = b.step_io(a)
            else:
                mk = (lambda t: t.pin_memory()) if flavour == "pinned" else (lambda t: t.cuda())
                ta = mk(torch.from_numpy(np.ascontiguousarray(a)).to(dt))
                tq = mk(torch.empty(nenv, model.nq, dtype=dt)); tv = mk(torch.empty(nenv, model.nv, dtype=dt))
                torch.cuda.synchronize()
                b.step_io_ptr(ta.data_ptr(), tq.data_ptr(), tv.data_ptr(), code, A.MEM_HOST if flavour == "pinned" else A.MEM_DEVICE)
                b.sync()
                q, v = tq.cpu().numpy(), tv.cpu().numpy()
            assert np.array_equal(q, wq) and np.array_equal(v, wv), flavour
        assert np.array_equal(b.get("ctrl"), ref.get("ctrl"))
