"""oxide_control_b200 - B200-native batched `mj_step` behind oxide_control's Physics API.

Only what the hot path needs lives here: csrc/ (CUDA kernels, MJCF compiler, C ABI -> lib/libox_b200.so),
the ctypes binding (_abi), the host-side mirror of the reference interface (physics) and the synthetic
models of BASELINE.json's configs (models). Importing the package loads the shared library and fails
loudly if it has not been built; there is no fallback implementation.
"""
from . import _abi
from .physics import (BatchedPhysics, PhysicsGroup, measure_fma_peak, Physics, Model, Actuators, ObjectId, obj, joint, Error, MujocoError, MjsError,
                      NameNotFound, PhysicsDiverged, JointTypeNotMatch, CudaError)
from .environment import Task, Observation, Action, Environment, TimeStep, TaskSpec, BatchedEnvironment, BatchedTimeStep
from . import models

_abi.lib()  # fail loudly at import time when the extension is missing

mjMAXVAL = 1e10
mjMINVAL = 1e-15
__all__ = ["BatchedPhysics", "PhysicsGroup", "measure_fma_peak", "Physics", "Model", "Actuators", "ObjectId", "obj", "joint", "Error", "MujocoError", "MjsError",
           "NameNotFound", "PhysicsDiverged", "JointTypeNotMatch", "CudaError", "models", "mjMAXVAL", "mjMINVAL",
           "Task", "Observation", "Action", "Environment", "TimeStep", "TaskSpec", "BatchedEnvironment", "BatchedTimeStep"]
