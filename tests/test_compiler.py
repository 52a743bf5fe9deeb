"""Model compiler known answers: mass properties vs closed forms, frames, addressing tables,
defaults, units, names and the error convention of Physics::from_xml_string (reference src/physics.rs:18-24, src/error.rs)."""
import os
import tempfile

import numpy as np
import pytest

import oxide_control_b200 as ox


def one_geom(geom, extra=""):
    return f'<mujoco><compiler angle="radian"/><worldbody><body name="b"><joint name="j"/><geom {geom}/>{extra}</body></worldbody></mujoco>'


def test_capsule_sphere_box_cylinder_ellipsoid_mass_and_inertia():
    r, l, rho = 0.07, 0.3, 1000.0
    m = ox.Model.from_xml_string(one_geom(f'type="capsule" size="{r} {l}"'))
    mc, mh = rho * np.pi * r * r * 2 * l, rho * 2 / 3 * np.pi * r ** 3
    assert np.isclose(m.body_mass[1], mc + 2 * mh, rtol=1e-14)
    Izz = 0.5 * mc * r * r + 2 * (2 / 5 * mh * r * r)
    Ixx = mc * (3 * r * r + 4 * l * l) / 12 + 2 * (mh * (83 / 320) * r * r + mh * (l + 3 * r / 8) ** 2)
    assert np.allclose(sorted(m.body_inertia[3:6]), sorted([Ixx, Ixx, Izz]), rtol=1e-13)
    m = ox.Model.from_xml_string(one_geom('type="sphere" size="0.2"'))
    ms = rho * 4 / 3 * np.pi * 0.2 ** 3
    assert np.isclose(m.body_mass[1], ms) and np.allclose(m.body_inertia[3:6], 0.4 * ms * 0.04)
    a, b, c = 0.1, 0.2, 0.3
    m = ox.Model.from_xml_string(one_geom(f'type="box" size="{a} {b} {c}"'))
    mb = rho * 8 * a * b * c
    assert np.isclose(m.body_mass[1], mb)
    assert np.allclose(sorted(m.body_inertia[3:6]), sorted([mb * (b * b + c * c) / 3, mb * (a * a + c * c) / 3, mb * (a * a + b * b) / 3]))
    m = ox.Model.from_xml_string(one_geom('type="cylinder" size="0.1 0.25"'))
    my = rho * np.pi * 0.01 * 0.5
    assert np.isclose(m.body_mass[1], my)
    assert np.allclose(sorted(m.body_inertia[3:6]), sorted([my * (3 * 0.01 + 0.25) / 12] * 2 + [my * 0.01 / 2]))
    m = ox.Model.from_xml_string(one_geom(f'type="ellipsoid" size="{a} {b} {c}"'))
    me = rho * 4 / 3 * np.pi * a * b * c
    assert np.allclose(sorted(m.body_inertia[3:6]), sorted([me * (b * b + c * c) / 5, me * (a * a + c * c) / 5, me * (a * a + b * b) / 5]))
    # explicit mass overrides density; planes are massless
    m = ox.Model.from_xml_string(one_geom('type="sphere" size="0.2" mass="3"'))
    assert m.body_mass[1] == 3.0


def test_two_geoms_combine_by_parallel_axis_and_principal_frame():
    xml = one_geom('type="sphere" size="0.1" mass="1" pos="1 0 0"', '<geom type="sphere" size="0.1" mass="3" pos="-1 0 0"/>')
    m = ox.Model.from_xml_string(xml)
    assert np.allclose(m.body_ipos[3:6], [-0.5, 0, 0]) and m.body_mass[1] == 4
    Is = 0.4 * 0.01
    axial, trans = Is * 4, Is * 4 + 1 * 1.5 ** 2 + 3 * 0.5 ** 2
    assert np.allclose(m.body_inertia[3:6], [trans, trans, axial])          # eigenvalues in decreasing order
    # principal frame: third axis (smallest inertia) along +-x
    from tests_util import quat2mat
    R = quat2mat(m.body_iquat[4:8])
    assert np.isclose(abs(R[0, 2]), 1) and np.isclose(np.linalg.det(R), 1)


def test_fromto_pose():
    m = ox.Model.from_xml_string(one_geom('type="capsule" fromto="0 0 0 0.6 0 0.8" size="0.05"'))
    assert np.allclose(m.geom_pos[:3], [0.3, 0, 0.4]) and np.isclose(m.geom_size[1], 0.5)
    from tests_util import quat2mat
    z = quat2mat(m.geom_quat[:4])[:, 2]
    assert np.allclose(np.abs(z @ np.array([0.6, 0, 0.8])), 1)


def test_inertial_element_overrides_geoms():
    xml = one_geom('type="sphere" size="0.2"', '<inertial pos="0 0 0.1" mass="2" diaginertia="0.3 0.2 0.1"/>')
    m = ox.Model.from_xml_string(xml)
    assert m.body_mass[1] == 2 and np.allclose(m.body_inertia[3:6], [0.3, 0.2, 0.1]) and np.allclose(m.body_ipos[3:6], [0, 0, 0.1])


def test_settotalmass_boundmass_boundinertia():
    """<compiler settotalmass> rescales every mass and inertia by one factor (dm_control's cheetah.xml sets 14 kg this way);
    boundmass / boundinertia are lower bounds applied to every body first."""
    body = ('<worldbody><body><joint/><geom type="sphere" size="0.1" density="1000"/><body pos="0 0 0.3"><joint/>'
            '<geom type="box" size="0.1 0.2 0.05" density="300"/></body></body></worldbody>')
    m0 = ox.Model.from_xml_string(f"<mujoco>{body}</mujoco>")
    m1 = ox.Model.from_xml_string(f'<mujoco><compiler settotalmass="14"/>{body}</mujoco>')
    k = 14 / m0.body_mass.sum()
    assert abs(m1.body_mass.sum() - 14) < 1e-12 and np.allclose(m1.body_mass, k * m0.body_mass, rtol=1e-13)
    assert np.allclose(m1.body_inertia, k * m0.body_inertia, rtol=1e-13) and np.allclose(m1.body_subtreemass[1], 14)
    assert np.allclose(m1.dof_invweight0, m0.dof_invweight0 / k, rtol=1e-12)             # everything downstream sees the new masses
    m2 = ox.Model.from_xml_string(f'<mujoco><compiler boundmass="5" boundinertia="0.02"/>{body}</mujoco>')
    assert np.allclose(m2.body_mass[1:], np.maximum(m0.body_mass[1:], 5)) and np.allclose(m2.body_inertia[3:], np.maximum(m0.body_inertia[3:], 0.02))
    m3 = ox.Model.from_xml_string(f'<mujoco><compiler boundmass="5" settotalmass="1"/>{body}</mujoco>')   # bounds first, then the scale
    assert np.allclose(m3.body_mass[1:], [0.5, 0.5])


def test_attributes_with_an_effect_are_never_silently_ignored():
    """Everything the schema accepts either acts on the step or is refused with Error::Mjs."""
    one = '<mujoco>{top}<worldbody><body {b}><joint name="j" {j}/><geom size="0.1" {g}/></body></worldbody>{tail}</mujoco>'
    cases = [dict(top='<option><flag sensor="disable"/></option>'), dict(top='<option><flag autoreset="disable"/></option>'),
             dict(top='<option><flag spring="disable"/></option>'), dict(top='<option><flag damper="disable"/></option>'),
             dict(top='<option><flag invdiscrete="enable"/></option>'), dict(top='<option actuatorgroupdisable="1"/>'),
             dict(top='<compiler inertiagrouprange="0 1"/>'), dict(top='<compiler balanceinertia="true"/>'),
             dict(top='<statistic meaninertia="2"/>'), dict(j='springdamper="0.1 1"'), dict(j='actuatorfrclimited="true"'),
             dict(j='actuatorgravcomp="true"'), dict(g='shellinertia="true"'),
             dict(j='range="-1 1"', tail='<actuator><position joint="j" inheritrange="1"/></actuator>'),
             dict(tail='<sensor><jointpos joint="j" cutoff="0.5"/></sensor>')]
    for c in cases:
        xml = one.format(top=c.get("top", ""), b=c.get("b", ""), j=c.get("j", ""), g=c.get("g", ""), tail=c.get("tail", ""))
        with pytest.raises(ox.MjsError, match="outside the supported subset"):
            ox.Model.from_xml_string(xml)
    ok = one.format(top='<option><flag midphase="disable" nativeccd="disable"/></option><statistic extent="2"/>', b="", j='springdamper="0 0"',
                    g='shellinertia="false"', tail='<sensor><jointpos joint="j" cutoff="0" noise="0.1"/></sensor>')
    assert ox.Model.from_xml_string(ok).nsensor == 1                                # no effect on the step in the subset: accepted


def test_topology_tables_of_the_benchmark_models():
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    assert (m.nq, m.nv, m.nu, m.nbody, m.ngeom, m.nM) == (9, 9, 6, 8, 9, 36)
    assert list(m.dof_parentid) == [-1, 0, 1, 2, 3, 4, 2, 6, 7]
    assert list(m.dof_Madr) == [0, 1, 3, 6, 10, 15, 21, 25, 30]
    assert list(m.body_parentid) == [0, 0, 1, 2, 3, 1, 5, 6] and list(m.body_rootid) == [0, 1, 1, 1, 1, 1, 1, 1]
    assert m.npair == 8 and m.nconmax == 16 and m.nefcmax == 12 + 64
    assert list(m.pair_geom1) == [0] * 8 and list(m.pair_geom2) == list(range(1, 9))   # plane first (lower type id)
    assert np.allclose(m.pair_friction[:5], [0.4, 0.4, 0.1, 0.1, 0.1])                 # elementwise max, unpacked to 5
    assert np.isclose(m.pair_solimp[0], 1e-4) and np.isclose(m.pair_solimp[1], 0.8)   # both geoms use the default class; d0=0 clamps to mjMINIMP
    assert np.isclose(m.body_subtreemass[1], m.body_mass[1:].sum())
    h = ox.Model.from_xml_string(ox.models.HUMANOID)
    assert (h.nq, h.nv, h.nu, h.nbody, h.nM) == (28, 27, 21, 14, 243)
    assert np.allclose(h.qpos0[:7], [0, 0, 1.5, 1, 0, 0, 0])
    assert h.jnt_type[0] == ox.joint.Free and list(h.dof_parentid[:6]) == [-1, 0, 1, 2, 3, 4]


def test_default_classes_childclass_and_units():
    xml = """<mujoco><default><joint damping="1"/><default class="leg"><joint damping="2" range="-30 60"/>
      <geom size="0.05"/></default></default>
      <worldbody><body childclass="leg"><joint name="a"/><geom/>
        <body pos="0 0 -0.3"><joint name="b" class="main"/><geom type="sphere" size="0.1"/></body></body></worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)  # default angle unit is DEGREE
    assert np.allclose(m.dof_damping, [2, 1])
    assert np.allclose(m.jnt_range[:2], np.deg2rad([-30, 60])) and m.jnt_limited[0] == 1 and m.jnt_limited[1] == 0
    assert np.isclose(m.geom_size[0], 0.05)
    e = ox.Model.from_xml_string('<mujoco><worldbody><body euler="90 0 0"><joint/><geom size="0.1"/></body></worldbody></mujoco>')
    assert np.allclose(e.body_quat[4:8], [np.sqrt(0.5), np.sqrt(0.5), 0, 0])


def test_actuator_shortcuts():
    xml = """<mujoco><compiler angle="radian"/><worldbody><body><joint name="j"/><geom size="0.1"/></body></worldbody>
      <actuator><motor joint="j" gear="3" ctrlrange="-2 2"/><position joint="j" kp="7" kv="0.5"/><velocity joint="j" kv="4" forcerange="-1 1"/>
      </actuator></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    assert list(m.actuator_gainprm.reshape(3, 3)[:, 0]) == [1, 7, 4]
    assert np.allclose(m.actuator_biasprm.reshape(3, 3), [[0, 0, 0], [0, -7, -0.5], [0, 0, -4]])
    assert list(m.actuator_biastype) == [0, 1, 1] and list(m.actuator_ctrllimited) == [1, 0, 0] and list(m.actuator_forcelimited) == [0, 0, 1]


def test_names_and_ids():
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    assert m.object_id(ox.obj.Body, "torso").index == 1 and m.object_id(ox.obj.Body, "world").index == 0
    assert m.object_id(ox.obj.Joint, "bfoot").index == 5 and m.object_name(ox.ObjectId(ox.obj.Actuator, 5)) == "ffoot"
    assert m.object_id(ox.obj.Geom, "floor").index == 0 and m.object_id(ox.obj.Sensor, "torso_subtreelinvel").index == 0
    assert m.object_id(ox.obj.Joint, "missing") is None and m.object_id(ox.obj.Plugin, "x") is None


def test_error_convention():
    with pytest.raises(ox.MujocoError):   # malformed XML -> Error::Mujoco
        ox.Model.from_xml_string("<mujoco><worldbody></mujoco>")
    with pytest.raises(ox.MujocoError):   # schema violation -> Error::Mujoco
        ox.Model.from_xml_string('<mujoco><worldbody><body><geom size="0.1" bogus="1"/></body></worldbody></mujoco>')
    with pytest.raises(ox.MujocoError):
        ox.Model.from_xml("/nonexistent/model.xml")
    for bad in ['<mujoco><option cone="elliptic" solver="PGS"/><worldbody/></mujoco>',
                '<mujoco><worldbody><body><joint range="1 -1"/><geom size="0.1"/></body></worldbody></mujoco>',
                '<mujoco><worldbody><body><joint/></body></worldbody></mujoco>',                       # massless moving body
                '<mujoco><worldbody><body><freejoint/><geom type="cylinder" size=".1 .1"/></body><body><freejoint/>'
                '<geom type="box" size=".1 .1 .1"/></body></worldbody></mujoco>',                    # cylinder collisions unsupported
                '<mujoco><worldbody><body name="a"><joint/><geom size="0.1"/></body><body name="a"><joint/><geom size="0.1"/>'
                '</body></worldbody></mujoco>']:
        with pytest.raises(ox.MjsError):  # compile failure -> Error::Mjs(message)
            ox.Model.from_xml_string(bad)
    with tempfile.NamedTemporaryFile("w", suffix=".xml", delete=False) as f:
        f.write(ox.models.PENDULUM)
    try:
        assert ox.Model.from_xml(f.name).nq == 1
    finally:
        os.unlink(f.name)
