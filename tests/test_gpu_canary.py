"""Guard-zone ("canary") checks for the GPU-only code that writes caller buffers: the pack kernels of bulk I/O, the fused
I/O of ox_batch_step_io, the observation / reward writers of the Environment layer and the state-record kernels.
compute-sanitizer is closed on the GPU pool (gpurun: "closed on this pool"), so out-of-bounds writes are hunted the way the
pool recommends: every device buffer handed to the library is surrounded by sentinel-filled guard zones that must survive,
and odd, non-multiple-of-32 batch sizes exercise the tail handling. The stage arithmetic itself runs under ASan/UBSan on
the host instantiation (tools/host_sanitizer_run.sh -> profiles/r2_host_asan_ubsan.log)."""
import ctypes as C

import numpy as np
import pytest
import torch

import oxide_control_b200 as ox
from oxide_control_b200 import _abi as A
from support import SEED, random_state

pytestmark = pytest.mark.gpu
GUARD = 4096  # elements of guard zone on each side
SENT = -12345.678


class Guarded:
    def __init__(self, n, dtype=torch.float32):
        self.n, self.dtype = n, dtype
        self.buf = torch.full((n + 2 * GUARD,), SENT if dtype.is_floating_point else 113, dtype=dtype, device="cuda")
        self.sent = self.buf[0].item()

    @property
    def ptr(self):
        return self.buf.data_ptr() + GUARD * self.buf.element_size()

    def body(self):
        return self.buf[GUARD:GUARD + self.n]

    def intact(self):
        torch.cuda.synchronize()
        return bool((self.buf[:GUARD] == self.sent).all() and (self.buf[GUARD + self.n:] == self.sent).all())


@pytest.mark.parametrize("name,nenv", [("cheetah", 97), ("humanoid", 45), ("cartpole", 1)])
def test_caller_buffers_are_written_within_bounds(name, nenv):
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    b = ox.BatchedPhysics(m, nenv, precision="f32")
    q, v = random_state(m, nenv, seed=71)
    b.set("qpos", q); b.set("qvel", v); b.ctrl_philox(True, SEED); b.step(30); b.sync()
    # bulk get, both layouts
    for field, cnt in (("qpos", m.nq), ("qvel", m.nv), ("qacc", m.nv)):
        for layout in (A.LAYOUT_ENV_MAJOR, A.LAYOUT_ELEM_MAJOR):
            g = Guarded(nenv * cnt)
            b.get_ptr(field, g.ptr, A.F32, A.MEM_DEVICE, layout)
            assert g.intact(), (field, layout)
            ref = b.get(field, np.float32, "env_major" if layout == A.LAYOUT_ENV_MAJOR else "elem_major")
            assert np.array_equal(g.body().cpu().numpy().reshape(ref.shape), ref)
    # fused step I/O with device buffers
    b.ctrl_philox(False, SEED)
    ctrl = Guarded(nenv * max(1, m.nu)); ctrl.body().uniform_(-1, 1)
    oq, ov = Guarded(nenv * m.nq), Guarded(nenv * m.nv)
    b.step_io_ptr(ctrl.ptr, oq.ptr, ov.ptr, A.F32, A.MEM_DEVICE); b.sync()
    assert ctrl.intact() and oq.intact() and ov.intact()
    assert np.array_equal(oq.body().cpu().numpy().reshape(nenv, m.nq), b.get("qpos", np.float32))
    # state record
    n = A.lib().ox_batch_state_size(b.handle)
    st = Guarded(nenv * n)
    assert A.lib().ox_batch_get_state(b.handle, C.c_void_p(st.ptr), A.F32, A.MEM_DEVICE) == 0
    assert st.intact() and torch.isfinite(st.body()).all()
    assert A.lib().ox_batch_set_state(b.handle, C.c_void_p(st.ptr), A.F32, A.MEM_DEVICE) == 0
    b.step(2); b.sync()
    assert np.isfinite(b.get("qpos")).all()


def test_environment_outputs_are_written_within_bounds():
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    nenv = 77
    b = ox.BatchedPhysics(m, nenv, precision="f32")
    env = ox.BatchedEnvironment(b, ox.TaskSpec(obs=[("qpos", 1, m.nq - 1), ("qvel", 0, m.nv)], reward=[("qvel", 0, "linear", 1.0)], time_limit=0.08,
                                               init_qpos_noise=0.1, init_qvel_noise=0.1, seed=5))
    act = Guarded(nenv * m.nu); act.body().uniform_(-1, 1)
    obs, rew, dis = Guarded(nenv * env.obs_dim), Guarded(nenv), Guarded(nenv)
    fin = Guarded(nenv, torch.uint8)
    for _ in range(12):     # crosses the time limit: auto-reset path included
        env.step_ptr(act.ptr, obs.ptr, rew.ptr, dis.ptr, fin.ptr, A.F32, A.MEM_DEVICE)
    b.sync()
    assert act.intact() and obs.intact() and rew.intact() and dis.intact() and fin.intact()
    assert torch.isfinite(obs.body()).all() and int(fin.body().sum()) >= 0
    env.close()
