"""N1: Environment / Task layer. CPU: the single-env mirror follows reference src/lib.rs:58-87 call for call;
the struct layouts of the ABI match the header. GPU (marked): `ox_env_*` (csrc/ox_env.cu) against a numpy restatement of
the same task evaluated on states produced by the CPU oracle."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oxide_control_b200 as ox
from oxide_control_b200 import _abi as A
from support import OracleData, philox4x32_10

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------ CPU: reference call order (src/lib.rs:58-87)
class _Log(list):
    pass


class _FakePhysics:
    def __init__(self, log): self.log, self.t = log, 0
    def actuators(self): self.log.append("actuators"); return self
    def step(self): self.log.append("physics.step"); self.t += 1


class _Obs(ox.Observation):
    def __init__(self, t): self.t = t
    @classmethod
    def generate(cls, physics): physics.log.append("generate"); return cls(physics.t)


class _Act(ox.Action):
    def apply(self, actuators): actuators.log.append("apply")


class _Task(ox.Task):
    Observation = _Obs
    def __init__(self, log): self.log = log
    def discount(self): self.log.append("discount"); return 0.99
    def init_episode(self, physics): self.log.append("init_episode"); physics.t = 0
    def should_finish_episode(self, observation): self.log.append("should_finish"); return observation.t >= 2
    def get_reward(self, observation, action): self.log.append("get_reward"); return float(observation.t)


def test_single_env_mirror_follows_reference_call_order():
    log = _Log()
    env = ox.Environment(_FakePhysics(log), _Task(log))
    obs = env.reset()
    assert log == ["init_episode", "generate"] and obs.t == 0          # src/lib.rs:58-61
    del log[:]
    ts = env.step(_Act())
    assert log == ["actuators", "apply", "physics.step", "generate", "get_reward", "should_finish", "discount"]  # :63-87
    assert not ts.is_finish and ts.reward == 1.0 and ts.discount == 0.99 and ts.observation.t == 1
    ts = env.step(_Act())
    assert ts.is_finish and ts.discount is None and ts.reward == 2.0    # TimeStep::Finish carries no discount (:55-59)
    assert env.task() is not None and env.physics() is env.physics_mut()


def test_task_spec_struct_layout_matches_header():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "ox_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(ox_obs_segment), sizeof(ox_reward_term), sizeof(ox_finish_cond), sizeof(ox_task_spec));
  printf("%zu %zu %zu %zu %zu %zu\n", offsetof(ox_task_spec, reward), offsetof(ox_task_spec, finish), offsetof(ox_task_spec, reward_bias),
         offsetof(ox_task_spec, seed), offsetof(ox_task_spec, frame_skip), offsetof(ox_reward_term, weight));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")], text=True).split()]
    T = A.TaskSpec
    want = [C.sizeof(A.ObsSegment), C.sizeof(A.RewardTerm), C.sizeof(A.FinishCond), C.sizeof(T), T.reward.offset, T.finish.offset,
            T.reward_bias.offset, T.seed.offset, T.frame_skip.offset, A.RewardTerm.weight.offset]
    assert got == want


def test_task_spec_defaults():
    s = A.TaskSpec()
    A.lib().ox_task_spec_default(C.byref(s))
    assert (s.discount, s.frame_skip, s.auto_reset, s.nobs, s.time_limit) == (1.0, 1, 1, 0, 0.0)
    assert A.lib().ox_env_obs_dim(None) == -1
    assert A.lib().ox_env_create(None, C.byref(s), C.byref(C.c_void_p())) == A.OX_ERR_INVALID


# ------------------------------------------------------------------ GPU
def _uniform(bits):
    return float(np.int32((int(bits) >> 9) * 2 + 1)) / 8388608.0 - 1.0


def _init_noise(m, gid, episode, seed, qa, va):
    """Restates env_init_episode (csrc/ox_env.cu): word w = Philox(ctr=(gid.lo, gid.hi, episode, w//4), key=seed^0x0E9150DE)[w%4]."""
    key = seed ^ 0x0E9150DE
    word = lambda w: _uniform(philox4x32_10((gid & 0xFFFFFFFF, gid >> 32, episode, w >> 2), (key & 0xFFFFFFFF, key >> 32))[w & 3])
    qpos = np.array(m.qpos0, dtype=np.float64).copy()
    for j in range(m.njnt):
        if int(m.jnt_type[j]) in (A.JNT_HINGE, A.JNT_SLIDE):
            a = int(m.jnt_qposadr[j])
            qpos[a] += qa * word(a)
    qvel = np.array([va * word(65536 + i) for i in range(m.nv)])
    return qpos, qvel


CHEETAH_TASK = dict(
    obs=[("qpos", 1, 8), ("qvel", 0, 9), ("sensordata", 0, 2)],
    reward=[("qvel", 0, "linear", 1.0)] + [("ctrl", i, "square", -0.1) for i in range(6)] + [("qpos", 2, "abs", -0.05)],
    finish=[("qpos", 2, -0.6, 0.6)], reward_bias=0.25, time_limit=0.25, discount=0.97, init_qpos_noise=0.1, init_qvel_noise=0.2, seed=77)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,frame_skip", [("f64", 1), ("f64", 3), ("f32", 1)])
def test_batched_environment_matches_numpy_task_on_oracle_states(precision, frame_skip):
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    nenv, nsteps, off = 48, 40, 1000
    b = ox.BatchedPhysics(m, nenv, precision=precision, env_id_offset=off)
    task = ox.TaskSpec(frame_skip=frame_skip, **CHEETAH_TASK)
    env = ox.BatchedEnvironment(b, task)
    assert env.obs_dim == 19
    rng = np.random.default_rng(3)
    tol = 1e-9 if precision == "f64" else 5e-3
    # oracle-side episodes
    ods, episode = [], np.zeros(nenv, int)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:], od.field("qvel")[:] = _init_noise(m, off + e, 0, task.seed, task.init_qpos_noise, task.init_qvel_noise)
        ods.append(od)
    obs0 = env.reset()
    for e, od in enumerate(ods):
        od.forward()
    want0 = np.stack([np.concatenate([od.field("qpos")[1:9], od.field("qvel"), od.field("sensordata")[:2]]) for od in ods])
    assert np.abs(obs0 - want0).max() <= tol
    nfin, ret_sum, len_sum = 0, 0.0, 0
    ep_ret, ep_len = np.zeros(nenv), np.zeros(nenv, int)
    for s in range(nsteps):
        act = rng.uniform(-1, 1, (nenv, m.nu))
        ts = env.step(act)
        for e, od in enumerate(ods):
            od.field("ctrl")[:] = act[e].astype(np.float32 if precision == "f32" else np.float64)
            for _ in range(frame_skip):
                od.step()
            obs = np.concatenate([od.field("qpos")[1:9], od.field("qvel"), od.field("sensordata")[:2]])
            r = 0.25 + od.field("qvel")[0] - 0.1 * float((od.field("ctrl") ** 2).sum()) - 0.05 * abs(od.field("qpos")[2])
            fin = not (-0.6 <= od.field("qpos")[2] <= 0.6) or od.field("time")[0] >= 0.25 - 0.5 * m.timestep
            assert np.abs(ts.observation[e] - obs).max() <= tol * max(1, s), (s, e)
            assert abs(ts.reward[e] - r) <= tol * max(1, s) * 10, (s, e)
            if precision == "f64":
                assert bool(ts.finished[e]) == fin, (s, e)
            fin = bool(ts.finished[e])
            assert ts.discount[e] == (0.0 if fin else np.float32(0.97) if precision == "f32" else 0.97)
            ep_ret[e] += ts.reward[e]; ep_len[e] += 1
            if fin:
                nfin += 1; ret_sum += ep_ret[e]; len_sum += ep_len[e]; ep_ret[e] = 0; ep_len[e] = 0
                episode[e] += 1
                od.reset()
                od.field("qpos")[:], od.field("qvel")[:] = _init_noise(m, off + e, int(episode[e]), task.seed, 0.1, 0.2)
            elif precision == "f32":   # keep the fp64 oracle on the fp32 trajectory so that thresholds are compared on equal states
                od.field("qpos")[:] = b.get1("qpos", e); od.field("qvel")[:] = b.get1("qvel", e)
                od.field("qacc_warmstart")[:] = b.get1("qacc_warmstart", e); od.field("time")[:] = b.get1("time", e)
        if s == nsteps - 1 or s == 12:   # auto-reset wrote exactly the restated init_episode state
            q = b.get("qpos")
            for e, od in enumerate(ods):
                if ep_len[e] == 0:
                    assert np.abs(q[e] - od.field("qpos")).max() <= 1e-6
    assert nfin >= nenv   # the 0.25 s time limit alone ends every episode at least once
    st = env.stats()
    assert st["episodes"] == nfin
    assert abs(st["mean_return"] - ret_sum / nfin) <= 1e-4 * max(1, abs(ret_sum / nfin)) and abs(st["mean_length"] - len_sum / nfin) < 1e-9


@pytest.mark.gpu
def test_env_device_buffers_no_auto_reset_and_errors():
    import torch
    m = ox.Model.from_xml_string(ox.models.CARTPOLE)
    nenv = 256
    b = ox.BatchedPhysics(m, nenv, precision="f32")
    with pytest.raises(ox.Error, match="out of range"):
        ox.BatchedEnvironment(b, ox.TaskSpec(obs=[("qpos", 0, 3)]))
    with pytest.raises(ox.Error, match="frame_skip"):
        ox.BatchedEnvironment(b, ox.TaskSpec(frame_skip=0))
    env = ox.BatchedEnvironment(b, ox.TaskSpec(obs=[("qpos", 0, 2), ("qvel", 0, 2)], reward=[("qpos", 1, "abs", -1.0)], time_limit=0.05,
                                               auto_reset=False, init_qpos_noise=0.05, seed=5))
    dev = torch.device("cuda:0")
    f32 = torch.float32   # explicit: tests/test_oracle_pins.py sets torch's default dtype to float64 at import time
    obs = torch.empty(nenv, 4, device=dev, dtype=f32); rew = torch.empty(nenv, device=dev, dtype=f32); dis = torch.empty(nenv, device=dev, dtype=f32)
    fin = torch.empty(nenv, dtype=torch.uint8, device=dev)
    act = torch.rand(nenv, 1, device=dev, dtype=f32) * 2 - 1
    torch.cuda.synchronize()
    env.reset_ptr(obs.data_ptr(), A.F32, A.MEM_DEVICE)
    b.sync()
    q0 = obs[:, :2].cpu().numpy().copy()
    assert 0 < np.abs(q0).max() <= 0.05 + 1e-6 and np.unique(q0[:, 1]).size > nenv // 2   # seeded noise differs per env
    nfin_steps = 0
    for s in range(8):
        env.step_ptr(act.data_ptr(), obs.data_ptr(), rew.data_ptr(), dis.data_ptr(), fin.data_ptr(), A.F32, A.MEM_DEVICE)
        b.sync()
        assert np.allclose(obs.cpu().numpy(), np.concatenate([b.get("qpos"), b.get("qvel")], 1))
        assert np.allclose(rew.cpu().numpy(), -np.abs(b.get("qpos")[:, 1]), atol=1e-7)
        f = fin.cpu().numpy().astype(bool)
        assert f.all() or not f.any()
        nfin_steps += int(f.all())
        assert np.array_equal(dis.cpu().numpy() == 0, f)
    assert nfin_steps >= 3                       # time limit 0.05 s: finishes and, without auto-reset, keeps finishing
    assert b.get("time").min() > 0.05            # no auto-reset: time keeps running
    # second reset draws a new episode index -> different noise
    env.reset_ptr(obs.data_ptr(), A.F32, A.MEM_DEVICE); b.sync()
    assert np.abs(obs[:, :2].cpu().numpy() - q0).max() > 1e-4
