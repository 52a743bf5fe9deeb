"""Environment / Task layer: the direct caller of the step.

Two mirrors of reference src/lib.rs:

* `Task`, `Observation`, `Action`, `Environment`, `TimeStep` - the reference's trait surface (src/lib.rs:8-88) for ONE
  `Physics`, with user-written Python bodies, call for call (`reset` = init_episode + generate; `step` = apply, physics.step,
  generate, get_reward, should_finish_episode -> Step{observation,reward,discount} | Finish{observation,reward}).
* `BatchedEnvironment` - the same contract for a whole `BatchedPhysics`, evaluated on the GPU by `csrc/ox_env.cu` through
  `ox_env_*` of the C ABI. User code cannot run inside a kernel, so the trait bodies become a declarative `TaskSpec`
  (observation = field slices, reward = bias + sum w*f(x), finish = range tests / time limit, init_episode = qpos0 + noise).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, Generic, List, Optional, Sequence, Tuple, TypeVar

import numpy as np

from . import _abi as A
from .physics import BatchedPhysics, Physics, _check

O = TypeVar("O")


# ------------------------------------------------------------------ single-env mirror of the reference traits
class Observation:
    """trait Observation (src/lib.rs:18-21): `generate(physics) -> Self`."""

    @classmethod
    def generate(cls, physics: Physics) -> "Observation":
        raise NotImplementedError


class Action:
    """trait Action (src/lib.rs:23-26): `apply(&self, actuators)`."""

    def apply(self, actuators) -> None:
        raise NotImplementedError


class Task:
    """trait Task (src/lib.rs:8-16)."""
    Observation: type = Observation

    def discount(self) -> float:
        raise NotImplementedError

    def init_episode(self, physics: Physics) -> None:
        raise NotImplementedError

    def should_finish_episode(self, observation) -> bool:
        raise NotImplementedError

    def get_reward(self, observation, action) -> float:
        raise NotImplementedError


@dataclass
class TimeStep(Generic[O]):
    """enum TimeStep<O> (src/lib.rs:50-60): `Step{observation,reward,discount}` | `Finish{observation,reward}`."""
    observation: O
    reward: float
    discount: Optional[float]  # None for Finish

    @property
    def is_finish(self) -> bool:
        return self.discount is None

    @staticmethod
    def Step(observation, reward, discount) -> "TimeStep":
        return TimeStep(observation, reward, discount)

    @staticmethod
    def Finish(observation, reward) -> "TimeStep":
        return TimeStep(observation, reward, None)


class Environment:
    """struct Environment<T: Task> (src/lib.rs:28-48, 62-88)."""

    def __init__(self, physics: Physics, task: Task):
        self._task, self._physics = task, physics

    def task(self) -> Task:
        return self._task

    def physics(self) -> Physics:
        return self._physics

    physics_mut = physics

    def reset(self):
        self._task.init_episode(self._physics)
        return self._task.Observation.generate(self._physics)

    def step(self, action: Action) -> TimeStep:
        action.apply(self._physics.actuators())
        self._physics.step()
        observation = self._task.Observation.generate(self._physics)
        reward = self._task.get_reward(observation, action)
        if self._task.should_finish_episode(observation):
            return TimeStep.Finish(observation, reward)
        return TimeStep.Step(observation, reward, self._task.discount())


# ------------------------------------------------------------------ batched, on device
@dataclass
class TaskSpec:
    """Declarative Task for BatchedEnvironment (ox_task_spec in include/ox_b200.h)."""
    obs: Sequence[Tuple[str, int, int]] = ()                 # (field, first, count): Observation::generate
    reward: Sequence[Tuple[str, int, str, float]] = ()       # (field, index, "linear"|"square"|"abs", weight): Task::get_reward
    finish: Sequence[Tuple[str, int, float, float]] = ()     # (field, index, lo, hi): Task::should_finish_episode
    reward_bias: float = 0.0
    time_limit: float = 0.0
    discount: float = 1.0
    init_qpos_noise: float = 0.0
    init_qvel_noise: float = 0.0
    seed: int = 0
    frame_skip: int = 1
    auto_reset: bool = True

    def to_c(self):
        kinds = {"linear": A.REWARD_LINEAR, "square": A.REWARD_SQUARE, "abs": A.REWARD_ABS}
        obs = (A.ObsSegment * max(1, len(self.obs)))(*[A.ObsSegment(A.FIELD[f], a, n) for f, a, n in self.obs])
        rew = (A.RewardTerm * max(1, len(self.reward)))(*[A.RewardTerm(A.FIELD[f], i, kinds[k], 0, w) for f, i, k, w in self.reward])
        fin = (A.FinishCond * max(1, len(self.finish)))(*[A.FinishCond(A.FIELD[f], i, lo, hi) for f, i, lo, hi in self.finish])
        s = A.TaskSpec()
        A.lib().ox_task_spec_default(C.byref(s))
        s.nobs, s.obs, s.nreward, s.reward, s.nfinish, s.finish = len(self.obs), obs, len(self.reward), rew, len(self.finish), fin
        s.reward_bias, s.time_limit, s.discount = self.reward_bias, self.time_limit, self.discount
        s.init_qpos_noise, s.init_qvel_noise, s.seed = self.init_qpos_noise, self.init_qvel_noise, self.seed
        s.frame_skip, s.auto_reset = self.frame_skip, int(self.auto_reset)
        return s, (obs, rew, fin)  # keep the arrays alive for the duration of the call


@dataclass
class BatchedTimeStep:
    """TimeStep for every env: `finished[e]` selects Finish (discount[e] = 0) or Step."""
    observation: np.ndarray  # [nenv, obs_dim]
    reward: np.ndarray       # [nenv]
    discount: np.ndarray     # [nenv]
    finished: np.ndarray     # [nenv] bool


class BatchedEnvironment:
    """Environment<T> over a BatchedPhysics, evaluated on device (csrc/ox_env.cu)."""

    def __init__(self, physics: BatchedPhysics, task: TaskSpec):
        self._physics, self._task = physics, task
        spec, keep = task.to_c()
        self._h = C.c_void_p()
        _check(A.lib().ox_env_create(physics.handle, C.byref(spec), C.byref(self._h)))
        del keep
        self.obs_dim = A.lib().ox_env_obs_dim(self._h)
        self.nenv = physics.nenv
        self._np = np.float64 if physics.precision == "f64" else np.float32
        self._code = A.F64 if physics.precision == "f64" else A.F32

    def close(self):
        if getattr(self, "_h", None):
            A.lib().ox_env_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def task(self) -> TaskSpec:
        return self._task

    def physics(self) -> BatchedPhysics:
        return self._physics

    @property
    def handle(self):
        return self._h

    def reset(self) -> np.ndarray:
        obs = np.empty((self.nenv, self.obs_dim), self._np)
        _check(A.lib().ox_env_reset(self._h, obs.ctypes.data_as(C.c_void_p), self._code, A.MEM_HOST))
        return obs

    def step(self, action: Optional[np.ndarray]) -> BatchedTimeStep:
        obs = np.empty((self.nenv, self.obs_dim), self._np)
        rew = np.empty(self.nenv, self._np)
        dis = np.empty(self.nenv, self._np)
        fin = np.empty(self.nenv, np.uint8)
        act = None
        if action is not None:
            act = np.ascontiguousarray(action, self._np)
            assert act.shape == (self.nenv, self._physics.model.nu), act.shape
        _check(A.lib().ox_env_step(self._h, act.ctypes.data_as(C.c_void_p) if act is not None else None, obs.ctypes.data_as(C.c_void_p),
                                   rew.ctypes.data_as(C.c_void_p), dis.ctypes.data_as(C.c_void_p), fin.ctypes.data_as(C.c_void_p),
                                   self._code, A.MEM_HOST))
        return BatchedTimeStep(obs, rew, dis, fin.astype(bool))

    # raw-pointer forms (pinned host or device buffers; what bench.py and a torch training loop use)
    def reset_ptr(self, obs_ptr: int, dtype_code: int, mem: int) -> None:
        _check(A.lib().ox_env_reset(self._h, C.c_void_p(obs_ptr), dtype_code, mem))

    def step_ptr(self, action_ptr: Optional[int], obs_ptr: Optional[int], reward_ptr: Optional[int], discount_ptr: Optional[int],
                 finished_ptr: Optional[int], dtype_code: int, mem: int) -> None:
        p = lambda x: C.c_void_p(x) if x else None
        _check(A.lib().ox_env_step(self._h, p(action_ptr), p(obs_ptr), p(reward_ptr), p(discount_ptr), p(finished_ptr), dtype_code, mem))

    def stats(self) -> dict:
        out = (C.c_double * 3)()
        _check(A.lib().ox_env_stats(self._h, out))
        n = out[0]
        return {"episodes": int(n), "mean_return": out[1] / n if n else float("nan"), "mean_length": out[2] / n if n else float("nan")}
