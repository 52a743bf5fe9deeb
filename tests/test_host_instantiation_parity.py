"""CPU stand-in for the GPU parity tests while no device is attached: the product's stage templates
(csrc/ox_stages.cuh, the code the CUDA kernels run) instantiated on the host by tests/native, against the oracle.
This checks the kernels' arithmetic and control flow, not the CUDA launch path; tests/test_gpu_parity.py does that."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err

FIELDS = ["xpos", "xquat", "xmat", "xipos", "ximat", "xanchor", "xaxis", "geom_xpos", "geom_xmat", "site_xpos", "subtree_com",
          "cinert", "cdof", "qM", "qLD", "qLDiagInv", "cvel", "cdof_dot", "qfrc_bias", "qfrc_passive", "actuator_force",
          "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qacc", "qfrc_constraint", "qacc_warmstart", "sensordata", "qpos", "qvel", "time"]


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "acrobot", "cheetah", "humanoid"])
def test_every_stage_output_matches_oracle_fp64(name):
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    nenv, nsteps = 6, 40
    qpos, qvel = random_state(m, nenv, seed=21)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        ods.append(od)
    for s in range(nsteps):
        for e, od in enumerate(ods):
            od.fill_ctrl_philox(e, s)
            od.step()
        hb.step(1, True, SEED, 0, s)
        for f in FIELDS:
            ref = np.stack([od.field(f) for od in ods])
            assert rel_err(hb.get(f), ref) <= 1e-9, (f, s)
        assert list(hb.get("ncon")[:, 0]) == [od.int("ncon") for od in ods]
        assert list(hb.get("nefc")[:, 0]) == [od.int("nefc") for od in ods]
        for e, od in enumerate(ods):  # contact and constraint lists row by row
            nc, ne = od.int("ncon"), od.int("nefc")
            assert np.array_equal(hb.get("con_pair")[e, :nc], od.int_field("con_pair"))
            assert rel_err(hb.get("con_dist")[e, :nc], od.field("con_dist")[:nc]) <= 1e-10
            assert rel_err(hb.get("con_frame")[e, :9 * nc], od.field("con_frame")[:9 * nc]) <= 1e-10
            assert rel_err(hb.get("efc_J")[e, :ne * m.nv], od.field("efc_J")[:ne * m.nv]) <= 1e-9
            for f in ("efc_pos", "efc_margin", "efc_D", "efc_aref", "efc_force"):
                assert rel_err(hb.get(f)[e, :ne], od.field(f)[:ne]) <= 1e-8, f


def test_fp32_instantiation_tracks_oracle_on_contact_free_models():
    for name, tol in (("pendulum", 1e-5), ("cartpole", 1e-4), ("acrobot", 1e-4)):
        m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
        qpos, qvel = random_state(m, 8, seed=22)
        hb = HostBatch(m, 8, "f32")
        hb.set("qpos", qpos); hb.set("qvel", qvel)
        hb.step(20, True, SEED, 0, 0)
        for e in range(8):
            od = OracleData(m)
            od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
            for s in range(20):
                od.fill_ctrl_philox(e, s)
                od.step()
            assert rel_err(hb.get("qpos")[e], od.field("qpos")) <= tol


def test_autoreset_on_nan_matches_oracle():
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    qpos, qvel = random_state(m, 4, seed=23)
    qvel[2, 3] = np.nan
    hb = HostBatch(m, 4, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(2)
    assert list(hb.get("diverged")[:, 0]) == [0, 0, 1, 0]
    od = OracleData(m)
    od.field("qpos")[:] = qpos[2]; od.field("qvel")[:] = qvel[2]
    od.step(); od.step()
    assert od.int("diverged") == 1 and rel_err(hb.get("qpos")[2], od.field("qpos")) <= 1e-12


@pytest.mark.parametrize("name", ["pendulum", "cartpole", "acrobot", "cheetah", "humanoid"])
@pytest.mark.parametrize("prec", ["f64", "f32"])
def test_model_specialised_step_is_bit_identical_to_generic(name, prec):
    """The generated specialisation (ox_specgen + csrc/ox_spec.cuh) runs the same stage code with compile-time tables
    and per-thread storage; on the host it must reproduce the generic path bit for bit."""
    import ctypes as C
    from support import hostcheck_lib
    L = hostcheck_lib()
    L.hc_step_spec.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64]
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    qpos, qvel = random_state(m, 5, seed=31)
    ha, hb = HostBatch(m, 5, prec), HostBatch(m, 5, prec)
    for h in (ha, hb):
        h.set("qpos", qpos); h.set("qvel", qvel)
    for s in range(0, 40, 4):
        ha.step(4, True, SEED, 10, s)
        assert L.hc_step_spec(hb.h, 4, 1, SEED, 10, s) == 1, "no specialisation compiled for this model"
        for f in ("qpos", "qvel", "qacc", "qacc_warmstart", "time", "sensordata", "ctrl", "ncon", "nefc", "solver_niter", "diverged"):
            assert np.array_equal(ha.get(f), hb.get(f)), (f, s)
