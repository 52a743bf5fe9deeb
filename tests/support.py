"""Test-side bindings: the CPU oracle (oracle/libox_oracle.so) and the host instantiation of the product's
stage templates (tests/native/libox_hostcheck.so). Both are checkers; neither is reachable from the product."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libox_oracle.so")
HOSTCHECK_SO = os.environ.get("OX_HOSTCHECK_SO") or os.path.join(ROOT, "tests", "native", "libox_hostcheck.so")   # the ASan build via tools/host_sanitizer_run.sh
if not os.path.isabs(HOSTCHECK_SO):
    HOSTCHECK_SO = os.path.join(ROOT, HOSTCHECK_SO)

SEED = 0x0B200


def _ensure(path: str, makedir: str) -> None:
    # always let make decide: the checkers share include/ox_b200.h (struct layout) with the product
    if os.path.exists(os.path.join(makedir, "Makefile")) and subprocess.call(["make", "-q", "-C", makedir], stdout=subprocess.DEVNULL,
                                                                              stderr=subprocess.DEVNULL) != 0:
        subprocess.check_call(["make", "-C", makedir], stdout=subprocess.DEVNULL)
    assert os.path.exists(path), path


_oracle = None


def oracle_lib() -> C.CDLL:
    global _oracle
    if _oracle is None:
        _ensure(ORACLE_SO, os.path.join(ROOT, "oracle"))
        L = C.CDLL(ORACLE_SO)
        P = C.c_void_p
        L.oxo_make_data.restype = P
        L.oxo_make_data.argtypes = [P]
        L.oxo_free_data.argtypes = [P]
        for f in ("oxo_reset", "oxo_forward", "oxo_step"):
            getattr(L, f).argtypes = [P, P]
        L.oxo_stage.argtypes = [P, P, C.c_char_p]
        L.oxo_field.restype = C.POINTER(C.c_double)
        L.oxo_field.argtypes = [P, C.c_char_p, C.POINTER(C.c_int32)]
        L.oxo_int.restype = C.c_int32
        L.oxo_int.argtypes = [P, C.c_char_p]
        L.oxo_int_field.restype = C.POINTER(C.c_int32)
        L.oxo_int_field.argtypes = [P, C.c_char_p, C.POINTER(C.c_int32)]
        L.oxo_fill_ctrl_philox.argtypes = [P, P, C.c_uint64, C.c_int64, C.c_int64]
        L.oxo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.oxo_bench.restype = C.c_double
        L.oxo_bench.argtypes = [P, C.c_int32, C.c_int32, C.c_int32, C.c_uint64, C.c_int64, C.c_int64, P, P, P]
        L.oxo_hardware_threads.restype = C.c_int32
        L.oxo_set_ctrl_scale.argtypes = [C.c_double]
        _oracle = L
    return _oracle


class OracleData:
    """One mjData-like env of the oracle."""

    def __init__(self, model):
        self.model = model
        self.L = oracle_lib()
        self.t = C.c_void_p(model.tables_ptr)
        self.d = C.c_void_p(self.L.oxo_make_data(self.t))

    def __del__(self):
        try:
            self.L.oxo_free_data(self.d)
        except Exception:
            pass

    def field(self, name: str) -> np.ndarray:
        """numpy view (no copy) of an AoS field."""
        n = C.c_int32()
        p = self.L.oxo_field(self.d, name.encode(), C.byref(n))
        if n.value < 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def int(self, name: str) -> int:
        return self.L.oxo_int(self.d, name.encode())

    def int_field(self, name: str) -> np.ndarray:
        n = C.c_int32()
        p = self.L.oxo_int_field(self.d, name.encode(), C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value > 0 else np.zeros(0, np.int32)

    def reset(self):
        self.L.oxo_reset(self.t, self.d)

    def forward(self):
        self.L.oxo_forward(self.t, self.d)

    def step(self):
        self.L.oxo_step(self.t, self.d)

    def stage(self, name: str):
        self.L.oxo_stage(self.t, self.d, name.encode())

    def fill_ctrl_philox(self, genv: int, stepno: int, seed: int = SEED):
        self.L.oxo_fill_ctrl_philox(self.t, self.d, seed, genv, stepno)


def oracle_bench(model, qpos, qvel, nsteps, nthreads, seed=SEED, env_off=0, step0=0):
    """Steps all envs on the CPU oracle; returns (seconds, stats[4]); qpos/qvel updated in place ([nenv][n], f64)."""
    L = oracle_lib()
    assert qpos.dtype == np.float64 and qvel.dtype == np.float64 and qpos.flags.c_contiguous and qvel.flags.c_contiguous
    stats = np.zeros(4)
    sec = L.oxo_bench(C.c_void_p(model.tables_ptr), qpos.shape[0], nsteps, nthreads, seed, env_off, step0,
                      qpos.ctypes.data_as(C.c_void_p), qvel.ctypes.data_as(C.c_void_p), stats.ctypes.data_as(C.c_void_p))
    return sec, stats


_hc = None


def hostcheck_lib() -> C.CDLL:
    global _hc
    if _hc is None:
        _ensure(HOSTCHECK_SO, os.path.join(ROOT, "tests", "native"))
        L = C.CDLL(HOSTCHECK_SO)
        P = C.c_void_p
        L.hc_create.restype = P
        L.hc_create.argtypes = [P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]
        L.hc_free.argtypes = [P]
        L.hc_stride.argtypes = [P]
        L.hc_reset.argtypes = [P]
        L.hc_forward.argtypes = [P]
        L.hc_step.argtypes = [P, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64]
        L.hc_field.restype = P
        L.hc_field.argtypes = [P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.hc_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        _hc = L
    return _hc


class HostBatch:
    """The product's stage templates run on the CPU (test-only), same SoA layout as the device arena."""

    def __init__(self, model, nenv, precision="f64", iterations=0, ls_iterations=0, tolerance=-1.0):
        from oxide_control_b200 import _abi as A
        self.A = A
        self.L = hostcheck_lib()
        self.nenv = nenv
        self.precision = precision
        self.h = C.c_void_p(self.L.hc_create(C.c_void_p(model.tables_ptr), nenv, 1 if precision == "f64" else 0,
                                             iterations, ls_iterations, tolerance))
        self.stride = self.L.hc_stride(self.h)
        self.L.hc_reset(self.h)

    def __del__(self):
        try:
            self.L.hc_free(self.h)
        except Exception:
            pass

    def _view(self, field: str) -> np.ndarray:
        n, isint = C.c_int(), C.c_int()
        p = self.L.hc_field(self.h, self.A.FIELD[field], C.byref(n), C.byref(isint))
        if n.value < 0:
            raise KeyError(field)
        if n.value == 0:
            return np.zeros((0, self.nenv))
        ct = C.c_int32 if isint.value else (C.c_double if self.precision == "f64" else C.c_float)
        arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n.value, self.stride))
        return arr[:, :self.nenv]

    def get(self, field: str) -> np.ndarray:
        """[nenv, n] copy (float64 for real fields)."""
        v = self._view(field).T
        return v.astype(np.float64) if v.dtype != np.int32 else v.copy()

    def set(self, field: str, values: np.ndarray):
        self._view(field)[:] = np.asarray(values).T

    def reset(self):
        self.L.hc_reset(self.h)

    def forward(self):
        self.L.hc_forward(self.h)

    def step(self, nsteps=1, philox=False, seed=SEED, env_off=0, step0=0):
        self.L.hc_step(self.h, nsteps, int(philox), seed, env_off, step0)


def random_state(model, nenv: int, seed: int = 0):
    """SURVEY 8d initial states: qpos0 + U(-0.1,0.1) on hinge/slide; free: z += U(0,0.2), quat = normalised(1,0,0,0 + N(0,0.05^2)),
    ball: same quaternion rule; qvel ~ N(0, 0.1^2)."""
    rng = np.random.default_rng(seed)
    nq, nv = model.nq, model.nv
    qpos = np.tile(np.asarray(model.qpos0, dtype=np.float64), (nenv, 1))
    for j in range(model.njnt):
        jt, a = int(model.jnt_type[j]), int(model.jnt_qposadr[j])
        if jt in (2, 3):
            qpos[:, a] += rng.uniform(-0.1, 0.1, nenv)
        else:
            if jt == 0:
                qpos[:, a + 2] += rng.uniform(0, 0.2, nenv)
                a += 3
            q = qpos[:, a:a + 4] + rng.normal(0, 0.05, (nenv, 4))
            qpos[:, a:a + 4] = q / np.linalg.norm(q, axis=1, keepdims=True)
    qvel = rng.normal(0, 0.1, (nenv, nv))
    return qpos, qvel


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max(1, |b|)  (SURVEY 8d parity gate formula)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))))


def philox4x32_10(ctr, key):
    """Pure-Python Philox4x32-10 (Random123), independent of both the oracle and the product; pinned by tests/test_philox.py."""
    c0, c1, c2, c3 = [int(x) & 0xFFFFFFFF for x in ctr]
    k0, k1 = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0, k1 = (k0 + 0x9E3779B9) & 0xFFFFFFFF, (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3
