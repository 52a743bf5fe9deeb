"""The C-ABI library loads without a GPU, exports every symbol include/ox_b200.h declares, and honours the error
convention on the paths that need no device (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import oxide_control_b200 as ox
from oxide_control_b200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ox_b200.h")).read()
    return sorted(set(re.findall(r"OX_API[^;(]*?\b(ox_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree_and_library_exports_everything():
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(A.SYMBOLS) == syms            # the ctypes "sys crate" declares exactly the header's entry points
    L = C.CDLL(A.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), s
    out = subprocess.check_output(["nm", "-D", "--defined-only", A.LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert set(syms) <= exported
    # nothing from the oracle or a host fallback is linked into the product
    assert not any(name.startswith("oxo_") or name.startswith("hc_") for name in exported)


def test_no_torch_or_oracle_dependency_in_the_shared_library():
    out = subprocess.check_output(["ldd", A.LIB_PATH], text=True)
    assert "torch" not in out and "ox_oracle" not in out and "hostcheck" not in out


def test_error_paths_without_device():
    L = A.lib()
    assert L.ox_version().startswith(b"ox_b200")
    h = C.c_void_p()
    assert L.ox_model_from_xml_string(None, C.byref(h)) == A.OX_ERR_INVALID
    assert L.ox_model_from_xml_string(b"<mujoco>", C.byref(h)) == A.OX_ERR_PARSE and b"line" in L.ox_last_error_message()
    assert L.ox_model_from_xml_path(b"/no/such/file.xml", C.byref(h)) == A.OX_ERR_IO
    assert L.ox_model_from_xml_string(ox.models.PENDULUM.encode(), C.byref(h)) == A.OX_OK
    assert L.ox_model_size(h, b"nq") == 1 and L.ox_model_size(h, b"bogus") == -1
    assert L.ox_model_name2id(h, A.OBJ_JOINT, b"hinge") == 0 and L.ox_model_name2id(h, A.OBJ_JOINT, b"x") == -1
    assert L.ox_model_id2name(h, A.OBJ_JOINT, 0) == b"hinge" and L.ox_model_id2name(h, A.OBJ_JOINT, 7) is None
    cfg = A.BatchConfig()
    L.ox_batch_config_default(C.byref(cfg))
    assert (cfg.nenv, cfg.precision, cfg.mode, cfg.tolerance) == (1, A.F32, A.MODE_FUSED, -1.0)
    b = C.c_void_p()
    assert L.ox_batch_create(None, C.byref(cfg), C.byref(b)) == A.OX_ERR_INVALID
    cfg.nenv = 0
    assert L.ox_batch_create(h, C.byref(cfg), C.byref(b)) == A.OX_ERR_INVALID
    for fn, args in ((L.ox_batch_step, (None, 1)), (L.ox_batch_forward, (None,)), (L.ox_batch_sync, (None,))):
        assert fn(*args) == A.OX_ERR_INVALID
    L.ox_model_free(h)


def test_batch_creation_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA the product refuses to create a batch."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    m = ox.Model.from_xml_string(ox.models.PENDULUM)
    with pytest.raises(ox.CudaError, match="no CPU fallback"):
        ox.BatchedPhysics(m, 4)
    with pytest.raises(ox.CudaError):
        ox.Physics.from_xml_string(ox.models.PENDULUM)
