"""N3 row of SURVEY 8f / VERDICT r1 next #8: sphere-box and capsule-box narrowphase (north_star names box contacts).

sphere-box follows mjc_SphereBox (clamp the centre to the box; a centre inside leaves through the nearest face).
capsule-box is this repo's own construction, NOT mjc_CapsuleBox (ORACLE_DECISIONS.md #12): sphere-box at the point of the axis
closest to the box - found exactly from the piecewise-linear derivative of the squared distance - plus the far end cap.
Closed-form geometry pins the oracle; tests/test_golden.py pins it on zoo_i against the dense checker, whose closest-point
search runs in exact rational arithmetic; host instantiation and (GPU) every kernel family must match the oracle."""
import numpy as np
import pytest

import oxide_control_b200 as ox
from support import HostBatch, OracleData, SEED, random_state, rel_err
from zoo_models import ZOO

SCENE = """<mujoco><compiler angle="radian"/><option gravity="0 0 0"/><worldbody>
<geom name="box" type="box" pos="0 0 0" size="0.5 0.3 0.2"/>
<body name="probe" pos="{pos}" euler="{euler}"><freejoint/><geom name="probe" type="{type}" size="{size}" margin="0.05"/></body>
</worldbody></mujoco>"""


def contacts(kind, size, pos, euler="0 0 0"):
    m = ox.Model.from_xml_string(SCENE.format(type=kind, size=size, pos=" ".join(map(str, pos)), euler=euler))
    od = OracleData(m)
    od.forward()
    n = od.int("ncon")
    return [(od.field("con_dist")[c], od.field("con_pos")[3 * c:3 * c + 3].copy(), od.field("con_frame")[9 * c:9 * c + 3].copy()) for c in range(n)]


def test_sphere_box_face_edge_corner_inside_and_miss():
    r = 0.1
    (d, p, n), = contacts("sphere", r, (0.1, 0.05, 0.28))                # above the +z face, 2 cm into it
    assert abs(d - (-0.02)) < 1e-15 and np.allclose(n, [0, 0, -1]) and np.allclose(p, [0.1, 0.05, 0.19])    # normal sphere -> box, point midway
    (d, p, n), = contacts("sphere", r, (0.56, 0.0, 0.26))                # off the +x / +z edge
    e = np.array([0.06, 0.0, 0.06])
    assert abs(d - (np.linalg.norm(e) - r)) < 1e-15 and np.allclose(n, -e / np.linalg.norm(e))
    (d, p, n), = contacts("sphere", r, (0.55, 0.35, 0.25))               # off the corner
    e = np.array([0.05, 0.05, 0.05])
    assert abs(d - (np.linalg.norm(e) - r)) < 1e-15 and np.allclose(n, -e / np.linalg.norm(e))
    (d, p, n), = contacts("sphere", r, (0.1, 0.27, 0.0))                 # centre INSIDE, nearest face +y (3 cm away)
    assert abs(d - (-0.03 - r)) < 1e-15 and np.allclose(n, [0, -1, 0])
    assert contacts("sphere", r, (0.1, 0.05, 0.36)) == []                # 6 cm gap > margin 0.05
    assert len(contacts("sphere", r, (0.1, 0.05, 0.34))) == 1            # 4 cm gap < margin: detected with positive distance


def test_capsule_box_flat_tilted_and_overhanging():
    size = "0.05 0.2"                                                     # radius, half length
    # lying flat along x on the top face, 1 cm into it: both end caps, same depth
    cs = contacts("capsule", size, (0.0, 0.0, 0.24), euler="0 90 0".replace("90", str(np.pi / 2)))
    assert len(cs) == 2 and all(abs(d - (-0.01)) < 1e-12 and np.allclose(n, [0, 0, -1], atol=1e-12) for d, _, n in cs)
    assert sorted(round(float(p[0]), 6) for _, p, _ in cs) == [-0.2, 0.2]
    # standing upright on the face: only the lower cap is within the margin
    (d, p, n), = contacts("capsule", size, (0.1, 0.1, 0.44))
    assert abs(d - (-0.01)) < 1e-12 and np.allclose(p[:2], [0.1, 0.1])
    # flat along x but hanging over the +x edge: the closest axis point is above the face, where the axis leaves the box's
    # footprint the distance starts to grow - the contact is at the lower (left) end and the far cap is out of reach
    cs = contacts("capsule", size, (0.55, 0.0, 0.24), euler=f"0 {np.pi / 2} 0")
    assert len(cs) >= 1 and abs(cs[0][0] - (-0.01)) < 1e-12
    # tilted: the lower end touches first
    ang = 0.3
    cs = contacts("capsule", size, (0.0, 0.0, 0.2 + 0.05 + 0.2 * np.sin(ang) - 0.005), euler=f"0 {np.pi / 2 - ang} 0")
    assert abs(cs[0][0] - (-0.005)) < 1e-9 and cs[0][1][0] < -0.15        # the axis is (cos a, 0, sin a): the -x end is the low one
    assert contacts("capsule", size, (0.0, 0.0, 0.6)) == []


BOXES = """<mujoco><compiler angle="radian"/><option gravity="0 0 0"/><worldbody>
<geom name="base" type="box" pos="0 0 0.2" size="0.5 0.4 0.2"/>
<body name="b" pos="{pos}" euler="{euler}"><freejoint/><geom name="b" type="box" size="0.1 0.1 0.1" margin="0.01"/></body></worldbody></mujoco>"""


def box_contacts(pos, euler="0 0 0"):
    m = ox.Model.from_xml_string(BOXES.format(pos=" ".join(map(str, pos)), euler=euler))
    od = OracleData(m)
    od.forward()
    return [(od.field("con_dist")[c], od.field("con_pos")[3 * c:3 * c + 3].copy(), od.field("con_frame")[9 * c:9 * c + 3].copy()) for c in range(od.int("ncon"))]


def test_box_box_known_answers():
    # face on face, 2 mm in: the four corners of the small box's bottom face, normal from geom1 (base) to geom2
    cs = box_contacts((0.1, 0.05, 0.498))
    assert len(cs) == 4 and all(abs(d + 0.002) < 1e-12 and np.allclose(n, [0, 0, 1]) for d, _, n in cs)
    assert sorted((round(p[0], 6), round(p[1], 6)) for _, p, _ in cs) == [(0.0, -0.05), (0.0, 0.15), (0.2, -0.05), (0.2, 0.15)]
    assert all(abs(p[2] - 0.399) < 1e-12 for _, p, _ in cs)                      # midway between the two surfaces
    # hanging over the +x edge of the base: the incident face is CLIPPED at x = 0.5
    cs = box_contacts((0.55, 0.0, 0.498))
    assert len(cs) == 4 and sorted(round(p[0], 6) for _, p, _ in cs) == [0.45, 0.45, 0.5, 0.5]
    # rotated 45 degrees about z: still four corners, on the diagonals
    cs = box_contacts((0.0, 0.0, 0.498), "0 0 0.7853981633974483")
    assert len(cs) == 4 and np.allclose(sorted(np.hypot(p[0], p[1]) for _, p, _ in cs), [0.1 * np.sqrt(2)] * 4)
    # standing on an edge (rotated about y): only the two vertices of the low edge are within the margin
    cs = box_contacts((0.0, 0.0, 0.53), "0 0.3 0")
    low = 0.53 - 0.1 * (np.cos(0.3) + np.sin(0.3))
    assert len(cs) == 2 and all(abs(d - (low - 0.4)) < 1e-12 for d, _, _ in cs)
    # on a corner: one contact straight below the centre
    cs = box_contacts((0.0, 0.0, 0.57), "0.6154797086703873 -0.7853981633974483 0")
    assert len(cs) == 1 and np.allclose(cs[0][1][:2], 0, atol=1e-9) and abs(cs[0][0] - (0.57 - 0.1 * np.sqrt(3) - 0.4)) < 1e-9
    assert box_contacts((0.0, 0.0, 0.8)) == []
    # two crossed edges (each box turned 45 degrees about a different axis): ONE contact, normal = the common perpendicular
    cs = box_contacts((0.5, 0.0, 0.53), "0.7853981633974483 0 0.7853981633974483")
    assert len(cs) == 1 and abs(np.linalg.norm(cs[0][2]) - 1) < 1e-12 and cs[0][2][2] > 0.5


def test_box_box_geometric_properties_on_random_poses():
    """Whatever the construction, a contact must be geometrically meaningful: unit normal along a separating-axis candidate, the
    reported distance equal to the separation measured along that normal (face contacts: vertex to face), the point within both
    boxes inflated by |dist| / 2 + margin, and never more than 8 contacts; separated boxes (beyond the margin) give none."""
    rng = np.random.default_rng(3)
    hits = 0
    for _ in range(300):
        pos = rng.uniform([-0.6, -0.5, 0.3], [0.6, 0.5, 0.62])
        eul = rng.uniform(-1.2, 1.2, 3)
        m = ox.Model.from_xml_string(BOXES.format(pos=" ".join(map(str, pos)), euler=" ".join(map(str, eul))))
        od = OracleData(m)
        od.forward()
        n = od.int("ncon")
        assert n <= 8
        R2 = od.field("geom_xmat")[9:18].reshape(3, 3)
        for c in range(n):
            dist, p, nrm = od.field("con_dist")[c], od.field("con_pos")[3 * c:3 * c + 3], od.field("con_frame")[9 * c:9 * c + 3]
            assert abs(np.linalg.norm(nrm) - 1) < 1e-12 and dist <= 0.01 + 1e-12
            slack = abs(dist) / 2 + 1e-9
            assert np.all(np.abs(p - np.array([0, 0, 0.2])) <= np.array([0.5, 0.4, 0.2]) + slack + 0.01)          # near / inside the base
            assert np.all(np.abs(R2.T @ (p - od.field("geom_xpos")[3:6])) <= 0.1 + slack + 0.01)                  # near / inside the small box
            hits += 1
    assert hits > 100


def test_pairs_are_listed_and_unsupported_shapes_refused():
    with pytest.raises(ox.MjsError, match="cylinder"):
        ox.Model.from_xml_string("<mujoco><worldbody><geom type='box' size='1 1 0.1'/><body pos='0 0 1'><freejoint/><geom type='cylinder' size='0.1 0.1'/></body></worldbody></mujoco>")
    m = ox.Model.from_xml_string(ZOO["zoo_i"])
    kinds = {(int(m.geom_type[a]), int(m.geom_type[b])): int(k) for a, b, k in zip(m.pair_geom1, m.pair_geom2, m.pair_maxcon)}
    assert kinds[(2, 6)] == 1 and kinds[(3, 6)] == 2 and kinds[(0, 6)] == 4
    p = ox.Model.from_xml_string(ZOO["zoo_p"])
    assert 8 in [int(k) for a, b, k in zip(p.pair_geom1, p.pair_geom2, p.pair_maxcon) if int(p.geom_type[a]) == 6 and int(p.geom_type[b]) == 6]


@pytest.mark.parametrize("nsteps", [400])
def test_box_stack_comes_to_rest_and_host_matches_oracle(nsteps):
    m = ox.Model.from_xml_string(ZOO["zoo_p"])
    nenv = 4
    qpos, qvel = random_state(m, nenv, seed=101)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.step()
        assert rel_err(hb.get("qpos")[e], od.field("qpos")) <= 1e-7 and hb.get("ncon")[e, 0] == od.int("ncon")
        z = od.field("qpos")[2]
        assert abs(z - (0.3 + 0.08)) < 5e-3, z            # the crate rests ON the table (table top 0.30, crate half height 0.08)
        assert np.abs(od.field("qvel")[:6]).max() < 5e-2


def test_host_instantiation_matches_oracle():
    m = ox.Model.from_xml_string(ZOO["zoo_i"])
    nenv, nsteps = 5, 250
    qpos, qvel = random_state(m, nenv, seed=53)
    hb = HostBatch(m, nenv, "f64")
    hb.set("qpos", qpos); hb.set("qvel", qvel)
    hb.step(nsteps, True, SEED, 0, 0)
    boxcon = 0
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        for s in range(nsteps):
            od.fill_ctrl_philox(e, s); od.step()
        for c in range(od.int("ncon")):
            p = od.int_field("con_pair")[c]
            boxcon += int(m.geom_type[m.pair_geom2[p]]) == 6 and int(m.geom_type[m.pair_geom1[p]]) in (2, 3)
        for f in ("qpos", "qvel", "qacc", "sensordata"):
            assert rel_err(hb.get(f)[e], od.field(f)) <= 1e-7, (f, e)
        assert hb.get("ncon")[e, 0] == od.int("ncon")
    assert boxcon >= 5            # spheres / capsules really rest on the boxes


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["zoo_i", "zoo_p"])
@pytest.mark.parametrize("mode,specialize", [("fused", 0), ("staged", 0), ("coop", 0), ("fused", 2)])
def test_gpu_vs_oracle(name, mode, specialize):
    m = ox.Model.from_xml_string(ZOO[name])
    nenv, nsteps = 64, 200
    qpos, qvel = random_state(m, nenv, seed=59)
    try:
        b = ox.BatchedPhysics(m, nenv, precision="f64", mode=mode, specialize=specialize)
    except ox.Error as err:
        assert mode == "coop" and "coop" in str(err) and name == "zoo_i"    # nv = 42 > 32: the cooperative kernel is not offered for that model
        return
    b.set("qpos", qpos); b.set("qvel", qvel); b.ctrl_philox(True, SEED)
    ods = []
    for e in range(nenv):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]; od.field("qvel")[:] = qvel[e]
        ods.append(od)
    done = 0
    for upto, tol in ((1, 1e-9), (120, 1e-6), (nsteps, 1e-4)):
        b.step(upto - done); b.sync()
        for e, od in enumerate(ods):
            for s in range(done, upto):
                od.fill_ctrl_philox(e, s); od.step()
        done = upto
        assert rel_err(b.get("qpos"), np.stack([od.field("qpos") for od in ods])) <= tol, upto
        assert np.array_equal(b.get("ncon")[:, 0], [od.int("ncon") for od in ods]) or upto > 120
    assert int(b.diverged().sum()) == 0
