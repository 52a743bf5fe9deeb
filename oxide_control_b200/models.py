"""Inline MJCF models for BASELINE.json's five configs.

pendulum, cartpole and acrobot are small models written for this repo. cheetah.xml and humanoid.xml are the public
dm_control / Gym assets (half-cheetah, humanoid) almost verbatim - same bodies, joints, geoms, actuators and options, minus
what the MJCF subset does not parse (textures, materials, cameras, tendons) - so that the benchmark configs are the models
the wider community steps with MuJoCo. The XML lives in spec_models/*.xml inside the package: the same files are read by
csrc/ox_specgen at build time to emit the compiled-in model-specialised step kernels (csrc/ox_spec.cuh); any other model is
specialised at run time (csrc/ox_jit.cpp).
"""
import os as _os


def _load(name: str) -> str:
    with open(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "spec_models", name + ".xml")) as f:
        return f.read()


# C1: single pendulum, 1 hinge, joint damping, no contacts. nq=nv=1, nu=1, nbody=2.
PENDULUM = _load("pendulum")

# C2: cartpole, slide + hinge, no contacts. nq=nv=2, nu=1, nbody=3, nM=3.
CARTPOLE = _load("cartpole")

# C3: acrobot / double pendulum, RK4, no damping, no contacts. nq=nv=2, nu=1, nbody=3.
ACROBOT = _load("acrobot")

# C4: half-cheetah-style planar chain, 8 capsules + plane, 6 limited+sprung+damped hinges, Newton/pyramidal.
# nq=nv=9, nu=6, nbody=8, ngeom=9, nM=36, ncon<=16, nefc<=12+64.
CHEETAH = _load("cheetah")

# C5: humanoid-style, free joint + 21 hinges, capsule/sphere geoms vs plane + a filtered set of self pairs.
# nq=28, nv=27, nu=21, nbody=14.
HUMANOID = _load("humanoid")

CONFIGS = {
    "pendulum": dict(xml=PENDULUM, nenv=1, precision="f64", label="C1 pendulum"),
    "cartpole": dict(xml=CARTPOLE, nenv=16384, precision="f32", label="C2 cartpole"),
    "acrobot": dict(xml=ACROBOT, nenv=65536, precision="f64", label="C3 acrobot RK4 fp64"),
    "cheetah": dict(xml=CHEETAH, nenv=8192, precision="f32", label="C4 cheetah"),
    "humanoid": dict(xml=HUMANOID, nenv=4096, precision="f32", label="C5 humanoid"),
}
