"""Binary model format (MuJoCo's mj_saveModel / mj_loadModel): a model written to disk and read back is
bit-identical to the one compiled from the XML - every table, option, size and name - and steps identically; corrupt,
truncated or foreign files are refused with an error, never misread."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from oxide_control_b200 import _abi as A
from support import OracleData, SEED, random_state
from zoo_models import HOPPER, NOCONTACT, ZOO

XML = {**{k: v["xml"] for k, v in ox.models.CONFIGS.items()}, **ZOO, **NOCONTACT, "hopper": HOPPER}
INT_TABLES = ["body_parentid", "body_mocapid", "jnt_type", "jnt_qposadr", "dof_Madr", "dof_Mdense", "geom_type", "pair_geom1", "pair_geom2",
              "actuator_trnid", "actuator_actadr", "sensor_type", "sensor_objid", "eq_type", "eq_obj1id", "eq_active0"]
REAL_TABLES = ["qpos0", "body_pos", "body_quat", "body_mass", "body_inertia", "body_invweight0", "dof_invweight0", "dof_damping", "jnt_range",
               "geom_size", "geom_friction", "pair_friction", "pair_solref", "actuator_gainprm", "actuator_biasprm", "eq_data", "eq_solref",
               "gravity"]
SIZES = ["nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "nsite", "nM", "npair", "nsensor", "nsensordata", "nconmax", "nefcmax", "nmocap", "neq",
         "integrator", "solver", "cone", "iterations", "ls_iterations", "disableflags"]


def _same(a, b):
    for s in SIZES:
        assert a.size(s) == b.size(s), s
    for t in INT_TABLES + REAL_TABLES:
        assert np.array_equal(getattr(a, t), getattr(b, t)), t
    for s in ("timestep", "tolerance", "ls_tolerance", "impratio", "meaninertia"):
        assert getattr(a, s) == getattr(b, s), s


@pytest.mark.parametrize("name", list(XML))
def test_round_trip_is_bit_identical(name, tmp_path):
    m = ox.Model.from_xml_string(XML[name])
    path = tmp_path / f"{name}.oxb"
    m.save(path)
    m2 = ox.Model.load(path)
    _same(m, m2)
    assert ox.Model.from_bytes(m.to_bytes()).to_bytes() == m.to_bytes() == path.read_bytes()
    # names survive (src/physics.rs:56-62)
    for objtype, count in ((ox.obj.Body, m.nbody), (ox.obj.Joint, m.njnt), (ox.obj.Actuator, m.nu), (ox.obj.Equality, m.neq)):
        for i in range(count):
            nm = m.object_name(ox.ObjectId(objtype, i))
            assert m2.object_name(ox.ObjectId(objtype, i)) == nm
            if nm:
                assert m2.object_id(objtype, nm).index == m.object_id(objtype, nm).index
    # and the loaded model steps exactly like the compiled one (CPU oracle on both)
    qpos, qvel = random_state(m, 1, seed=3)
    a, b = OracleData(m), OracleData(m2)
    for od in (a, b):
        od.field("qpos")[:] = qpos[0]; od.field("qvel")[:] = qvel[0]
    for s in range(30):
        a.fill_ctrl_philox(0, s); b.fill_ctrl_philox(0, s)
        a.step(); b.step()
    assert np.array_equal(a.field("qpos"), b.field("qpos")) and np.array_equal(a.field("qacc"), b.field("qacc"))


def test_bad_files_are_refused(tmp_path):
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    blob = bytearray(m.to_bytes())
    with pytest.raises(ox.Error, match="could not open"):
        ox.Model.load(tmp_path / "missing.oxb")
    with pytest.raises(ox.Error, match="bad magic"):
        ox.Model.from_bytes(b"MJB?" + bytes(blob[4:]))
    with pytest.raises(ox.Error, match="checksum|truncated"):
        ox.Model.from_bytes(bytes(blob[:len(blob) // 2]))
    flipped = bytearray(blob); flipped[len(blob) // 2] ^= 0x40
    with pytest.raises(ox.Error, match="checksum"):
        ox.Model.from_bytes(bytes(flipped))
    ver = bytearray(blob); ver[8] = 99
    with pytest.raises(ox.Error, match="version"):
        ox.Model.from_bytes(bytes(ver))
    (tmp_path / "text.oxb").write_text("<mujoco/>")
    with pytest.raises(ox.Error):
        ox.Model.load(tmp_path / "text.oxb")
    assert A.lib().ox_model_serialize(None, None, 0) == -1


@pytest.mark.gpu
def test_loaded_model_selects_the_same_specialised_kernel_and_steps_identically(tmp_path):
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    m.save(tmp_path / "cheetah.oxb")
    m2 = ox.Model.load(tmp_path / "cheetah.oxb")
    qpos, qvel = random_state(m, 256, seed=9)
    out = []
    for mm in (m, m2):
        b = ox.BatchedPhysics(mm, 256, precision="f32")
        assert b.kernel_name() == "cheetah"            # same model hash -> the compiled-in specialisation
        b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
        b.step(50); b.sync()
        out.append((b.get("qpos"), b.get("qacc")))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
