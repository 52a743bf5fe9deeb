"""What pins the oracle in lieu of reference golden vectors.

The reference holds no tests or fixtures for mj_step and libmujoco is absent (parity unpinned), so the
restated oracle is checked here against closed forms and against algorithm-independent derivations:
an autograd Lagrangian (mass matrix, bias forces), scipy's minimiser on the same convex objective,
KKT conditions, and structural invariants.
"""
import numpy as np
import pytest
import torch

import oxide_control_b200 as ox
from support import OracleData, random_state

torch.set_default_dtype(torch.float64)


def dense_M(model, qM):
    nv = model.nv
    M = np.zeros((nv, nv))
    for i in range(nv):
        adr, j = int(model.dof_Madr[i]), i
        while j >= 0:
            M[i, j] = M[j, i] = qM[adr]
            adr += 1
            j = int(model.dof_parentid[j])
    return M


# ---------------------------------------------------------------- D.2 pendulum closed form
def test_pendulum_acceleration_and_euler_step_closed_form():
    m = ox.Model.from_xml_string(ox.models.PENDULUM)
    mass, lc = m.body_mass[1], 0.25
    I_pivot = m.body_inertia[3] + mass * lc * lc
    g, b, h = 9.81, 0.1, m.timestep
    od = OracleData(m)
    th, om, u = 0.7, -0.4, 0.35
    od.field("qpos")[0], od.field("qvel")[0], od.field("ctrl")[0] = th, om, u
    od.forward()
    # hinge about +y, pole hanging along -z: gravity torque = -m g lc sin(theta)
    qacc = (-mass * g * lc * np.sin(th) - b * om + u) / I_pivot
    assert abs(od.field("qacc")[0] - qacc) < 1e-12
    assert abs(m.meaninertia - I_pivot) < 1e-15
    od.step()
    # implicit-in-velocity damping: (I + h b) qacc' = I qacc
    om1 = om + h * qacc * I_pivot / (I_pivot + h * b)
    assert abs(od.field("qvel")[0] - om1) < 1e-13
    assert abs(od.field("qpos")[0] - (th + h * om1)) < 1e-13
    assert abs(od.field("time")[0] - h) < 1e-16


def test_pendulum_1000_steps_vs_independent_scalar_integrator():
    m = ox.Model.from_xml_string(ox.models.PENDULUM)
    mass, lc, g, b, h = m.body_mass[1], 0.25, 9.81, 0.1, m.timestep
    I = m.body_inertia[3] + mass * lc * lc
    od = OracleData(m)
    th, om = 1.0, 0.0
    od.field("qpos")[0] = th
    for s in range(1000):
        u = 0.5 * np.sin(0.01 * s)
        od.field("ctrl")[0] = u
        od.step()
        qacc = (-mass * g * lc * np.sin(th) - b * om + u) / (I + h * b)
        om += h * qacc
        th += h * om
    assert abs(od.field("qpos")[0] - th) < 1e-11 and abs(od.field("qvel")[0] - om) < 1e-11


# ---------------------------------------------------------------- D.3 cartpole textbook equations of motion
def test_cartpole_mass_matrix_and_bias_closed_form():
    m = ox.Model.from_xml_string(ox.models.CARTPOLE)
    mc, mp, l = m.body_mass[1], m.body_mass[2], 0.5
    Ip = m.body_inertia[3 * 2]  # transverse inertia of the pole about its com (largest eigenvalue first)
    rng = np.random.default_rng(0)
    for _ in range(10):
        x, th, xd, thd = rng.normal(size=4)
        od = OracleData(m)
        od.field("qpos")[:] = [x, th]
        od.field("qvel")[:] = [xd, thd]
        od.forward()
        M = dense_M(m, od.field("qM"))
        # pole com at (x + l sin th, 1 + l cos th) for a hinge about +y
        Mref = np.array([[mc + mp, mp * l * np.cos(th)], [mp * l * np.cos(th), Ip + mp * l * l]])
        assert np.allclose(M, Mref, atol=1e-13)
        bias = np.array([-mp * l * np.sin(th) * thd ** 2, -mp * 9.81 * l * np.sin(th)])
        assert np.allclose(od.field("qfrc_bias"), bias, atol=1e-12)


# ---------------------------------------------------------------- D.4 autograd Lagrangian (algorithm independent)
def _quat2mat(q):
    w, x, y, z = q
    return torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
                        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
                        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])])


def _hat(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def _rodrigues(axis, ang):
    K = _hat(axis)
    return torch.eye(3) + torch.sin(ang) * K + (1 - torch.cos(ang)) * (K @ K)


def fk_com(model, qpos_base, x):
    """Forward kinematics written only from the model tables. x are nv generalised coordinates: hinge/slide
    coordinates themselves, and for free/ball joints a local body-frame rotation vector (evaluated at 0)
    on top of the base quaternion. Returns com positions [nbody,3] and inertial-frame rotations [nbody,3,3]."""
    T = lambda a: torch.tensor(np.asarray(a, dtype=np.float64))
    nb = model.nbody
    pos, R = [torch.zeros(3)], [torch.eye(3)]
    for i in range(1, nb):
        p = int(model.body_parentid[i])
        bp, bq = T(model.body_pos[3 * i:3 * i + 3]), T(model.body_quat[4 * i:4 * i + 4])
        P = pos[p] + R[p] @ bp
        Rb = R[p] @ _quat2mat(bq)
        for j in range(int(model.body_jntadr[i]), int(model.body_jntadr[i]) + int(model.body_jntnum[i])):
            jt, qa, da = int(model.jnt_type[j]), int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
            ax, jp = T(model.jnt_axis[3 * j:3 * j + 3]), T(model.jnt_pos[3 * j:3 * j + 3])
            if jt == 2:
                P = P + Rb @ ax * (x[da] - model.qpos0[qa])
            elif jt == 3:
                anchor = P + Rb @ jp
                Rb = Rb @ _rodrigues(ax, x[da] - model.qpos0[qa])
                P = anchor - Rb @ jp
            elif jt == 0:
                P = x[da:da + 3]
                th = x[da + 3:da + 6]
                Rb = _quat2mat(T(qpos_base[qa + 3:qa + 7])) @ (torch.eye(3) + _hat(th) + 0.5 * _hat(th) @ _hat(th))
            else:
                anchor = P + Rb @ jp
                th = x[da:da + 3]
                Rb = Rb @ _quat2mat(T(qpos_base[qa:qa + 4])) @ (torch.eye(3) + _hat(th) + 0.5 * _hat(th) @ _hat(th))
                P = anchor - Rb @ jp
        pos.append(P)
        R.append(Rb)
    com = torch.stack([pos[i] + R[i] @ T(model.body_ipos[3 * i:3 * i + 3]) for i in range(nb)])
    Ri = torch.stack([R[i] @ _quat2mat(T(model.body_iquat[4 * i:4 * i + 4])) for i in range(nb)])
    return com, Ri


def coords_from_qpos(model, qpos):
    x = np.zeros(model.nv)
    for j in range(model.njnt):
        jt, qa, da = int(model.jnt_type[j]), int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
        if jt in (2, 3):
            x[da] = qpos[qa]
        elif jt == 0:
            x[da:da + 3] = qpos[qa:qa + 3]
    return torch.tensor(x)


def kinetic_energy(model, qpos_base, x, xd):
    from torch.autograd.functional import jacobian
    com, Ri = fk_com(model, qpos_base, x)
    Jc = jacobian(lambda y: fk_com(model, qpos_base, y)[0], x, create_graph=True)    # [nb,3,nv]
    JR = jacobian(lambda y: fk_com(model, qpos_base, y)[1], x, create_graph=True)    # [nb,3,3,nv]
    T = torch.zeros(())
    for i in range(1, model.nbody):
        v = Jc[i] @ xd
        Rdot = JR[i] @ xd
        W = Rdot @ Ri[i].T
        w = torch.stack([W[2, 1], W[0, 2], W[1, 0]])
        I = Ri[i] @ torch.diag(torch.tensor(model.body_inertia[3 * i:3 * i + 3])) @ Ri[i].T
        T = T + 0.5 * model.body_mass[i] * (v @ v) + 0.5 * (w @ I @ w)
    return T


def lagrangian_M_bias(model, qpos, qvel, want_bias=True):
    from torch.autograd.functional import hessian, jacobian
    x0, xd0 = coords_from_qpos(model, qpos), torch.tensor(np.asarray(qvel, dtype=np.float64))
    M = hessian(lambda xd: kinetic_energy(model, qpos, x0, xd), xd0).numpy()
    if not want_bias:
        return M, None
    grav = torch.tensor(model.gravity)
    V = lambda x: -sum(model.body_mass[i] * (grav @ fk_com(model, qpos, x)[0][i]) for i in range(1, model.nbody))
    p = lambda x: jacobian(lambda xd: kinetic_energy(model, qpos, x, xd), xd0, create_graph=True)  # dT/dxd as f(x)
    dp_dx = jacobian(p, x0)                                                                           # [nv,nv]
    dT_dx = jacobian(lambda x: kinetic_energy(model, qpos, x, xd0), x0)
    dV_dx = jacobian(V, x0)
    bias = dp_dx @ xd0 - dT_dx + dV_dx
    return M, bias.numpy()


@pytest.mark.parametrize("name", ["cartpole", "acrobot", "cheetah"])
def test_mass_matrix_and_bias_vs_autograd_lagrangian(name):
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    qpos, qvel = random_state(m, 2, seed=5)
    qvel *= 10  # make Coriolis terms matter
    for e in range(2):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]
        od.field("qvel")[:] = qvel[e]
        od.forward()
        M, bias = lagrangian_M_bias(m, qpos[e], qvel[e])
        Mo = dense_M(m, od.field("qM")) - np.diag(m.dof_armature)
        assert np.allclose(Mo, M, atol=1e-11, rtol=1e-11), np.abs(Mo - M).max()
        assert np.allclose(od.field("qfrc_bias"), bias, atol=1e-10, rtol=1e-10), np.abs(od.field("qfrc_bias") - bias).max()


def test_humanoid_mass_matrix_vs_autograd_lagrangian():
    """Free joint + 21 hinges: M only (quasi-velocities make the Lagrange bias form inapplicable)."""
    m = ox.Model.from_xml_string(ox.models.HUMANOID)
    qpos, qvel = random_state(m, 1, seed=6)
    od = OracleData(m)
    od.field("qpos")[:] = qpos[0]
    od.forward()
    M, _ = lagrangian_M_bias(m, qpos[0], np.zeros(m.nv), want_bias=False)
    Mo = dense_M(m, od.field("qM")) - np.diag(m.dof_armature)
    assert np.allclose(Mo, M, atol=1e-10, rtol=1e-10), np.abs(Mo - M).max()


def test_compiler_invweight0_vs_autograd_mass_matrix():
    """dof_invweight0 / meaninertia come from the compiler's dense path; check them against the Lagrangian M at qpos0."""
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    M, _ = lagrangian_M_bias(m, np.asarray(m.qpos0), np.zeros(m.nv), want_bias=False)
    M = M + np.diag(m.dof_armature)
    assert abs(m.meaninertia - np.mean(np.diag(M))) < 1e-11
    assert np.allclose(m.dof_invweight0, np.diag(np.linalg.inv(M)), rtol=1e-10)


# ---------------------------------------------------------------- D.5 structural invariants
@pytest.mark.parametrize("name", ["cheetah", "humanoid"])
def test_factorisation_and_inverse_dynamics_invariants(name):
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    qpos, qvel = random_state(m, 1, seed=7)
    od = OracleData(m)
    od.field("qpos")[:] = qpos[0]
    od.field("qvel")[:] = qvel[0]
    od.fill_ctrl_philox(0, 0)
    for _ in range(60):  # fall into contact
        od.step()
    od.forward()
    nv = m.nv
    M = dense_M(m, od.field("qM"))
    assert np.allclose(M, M.T) and np.all(np.linalg.eigvalsh(M) > 0)
    # L' D L reconstructs M: qLD holds D on the diagonal and L(i, ancestors) off it
    L, D = np.eye(nv), np.zeros(nv)
    qLD = od.field("qLD")
    for i in range(nv):
        adr, j = int(m.dof_Madr[i]), int(m.dof_parentid[i])
        D[i] = qLD[adr]
        adr += 1
        while j >= 0:
            L[i, j] = qLD[adr]
            adr += 1
            j = int(m.dof_parentid[j])
    assert np.allclose(L.T @ np.diag(D) @ L, M, atol=1e-12 * np.abs(M).max())
    assert np.allclose(od.field("qLDiagInv"), 1 / D)
    # qacc_smooth = M^-1 qfrc_smooth
    assert np.allclose(M @ od.field("qacc_smooth"), od.field("qfrc_smooth"), atol=1e-9)
    # inverse dynamics residual with the constraint force
    res = M @ od.field("qacc") + od.field("qfrc_bias") - od.field("qfrc_passive") - od.field("qfrc_actuator") \
        - od.field("qfrc_applied") - od.field("qfrc_constraint")
    assert np.abs(res).max() < 1e-7 * max(1.0, np.abs(od.field("qfrc_bias")).max())
    assert od.int("nefc") > 0


def test_free_fall_and_gravity_bias():
    xml = """<mujoco><compiler angle="radian"/><worldbody><body pos="0 0 3"><freejoint/>
             <geom type="box" size="0.1 0.2 0.3" contype="0" conaffinity="0"/></body></worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m)
    od.field("qvel")[:] = [0.3, -0.2, 0.1, 0.0, 0.0, 0.0]
    od.forward()
    assert np.allclose(od.field("qacc"), [0, 0, -9.81, 0, 0, 0], atol=1e-12)
    # torque-free spinning box under RK4 conserves angular momentum (world frame) and keeps a unit quaternion
    xml4 = xml.replace('<worldbody>', '<option integrator="RK4" gravity="0 0 0"/><worldbody>')
    m4 = ox.Model.from_xml_string(xml4)
    od = OracleData(m4)
    od.field("qvel")[3:] = [1.0, 2.0, 0.5]
    def ang_mom():
        od.forward()
        R = od.field("ximat")[9:18].reshape(3, 3)
        I = R @ np.diag(m4.body_inertia[3:6]) @ R.T
        Rb = od.field("xmat")[9:18].reshape(3, 3)
        return I @ (Rb @ od.field("qvel")[3:6])
    L0 = ang_mom()
    for _ in range(500):
        od.step()
    assert np.allclose(ang_mom(), L0, rtol=1e-5)  # 1.4e-6 observed (qpos integrates on SO(3) with the RK4-averaged velocity)
    assert abs(np.linalg.norm(od.field("qpos")[3:7]) - 1) < 1e-12


# ---------------------------------------------------------------- D.7 constraint solver
def _objective(m, od):
    nv, nefc = m.nv, od.int("nefc")
    M = dense_M(m, od.field("qM"))
    J = od.field("efc_J")[:nefc * nv].reshape(nefc, nv).copy()
    D, aref = od.field("efc_D")[:nefc].copy(), od.field("efc_aref")[:nefc].copy()
    a0, f0 = od.field("qacc_smooth").copy(), od.field("qfrc_smooth").copy()
    def cost(a):
        jar = J @ a - aref
        act = jar < 0
        return 0.5 * (a - a0) @ M @ (a - a0) + 0.5 * np.sum(D[act] * jar[act] ** 2)
    def grad(a):
        jar = J @ a - aref
        f = np.where(jar < 0, -D * jar, 0.0)
        return M @ a - f0 - J.T @ f
    return cost, grad, J, D, aref


@pytest.mark.parametrize("name", ["cheetah", "humanoid"])
def test_solver_kkt_and_scipy_agreement(name):
    from scipy.optimize import minimize
    m = ox.Model.from_xml_string(ox.models.CONFIGS[name]["xml"])
    qpos, qvel = random_state(m, 3, seed=8)
    for e in range(3):
        od = OracleData(m)
        od.field("qpos")[:] = qpos[e]
        od.field("qvel")[:] = qvel[e]
        for s in range(50 + 10 * e):
            od.fill_ctrl_philox(e, s)
            od.step()
        od.forward()
        assert od.int("nefc") > 0
        cost, grad, J, D, aref = _objective(m, od)
        a = od.field("qacc").copy()
        scale = 1.0 / (m.meaninertia * max(1, m.nv))
        assert scale * np.linalg.norm(grad(a)) < 1e-7                     # stationarity
        f = od.field("efc_force")[:od.int("nefc")]
        jar = J @ a - aref
        assert np.all(f >= 0) and np.abs(f * np.maximum(0, jar)).max() < 1e-9   # complementarity
        assert np.allclose(od.field("qfrc_constraint"), J.T @ f, atol=1e-9)
        r = minimize(cost, od.field("qacc_smooth"), jac=grad, method="BFGS", options={"gtol": 1e-10, "maxiter": 5000})
        assert cost(a) <= r.fun + 1e-9 * max(1, abs(r.fun))
        assert np.allclose(a, r.x, atol=1e-5 * max(1, np.abs(a).max()))


def test_cg_solver_matches_newton():
    xml_cg = ox.models.CHEETAH.replace('<option timestep="0.01"/>', '<option timestep="0.01" solver="CG" tolerance="1e-12"/>')
    xml_nt = ox.models.CHEETAH.replace('<option timestep="0.01"/>', '<option timestep="0.01" tolerance="1e-12"/>')
    mc, mn = ox.Model.from_xml_string(xml_cg), ox.Model.from_xml_string(xml_nt)
    qpos, qvel = random_state(mn, 1, seed=9)
    oc, on = OracleData(mc), OracleData(mn)
    for od in (oc, on):
        od.field("qpos")[:] = qpos[0]
        od.field("qvel")[:] = qvel[0]
    for s in range(60):
        for od in (oc, on):
            od.fill_ctrl_philox(0, s)
            od.step()
    on.forward()
    oc.field("qpos")[:] = on.field("qpos"); oc.field("qvel")[:] = on.field("qvel"); oc.field("qacc_warmstart")[:] = 0
    on.field("qacc_warmstart")[:] = 0
    on.forward(); oc.forward()
    assert on.int("nefc") > 0
    assert np.allclose(oc.field("qacc"), on.field("qacc"), atol=1e-5 * max(1, np.abs(on.field("qacc")).max()))


def test_resting_sphere_equilibrium_penetration():
    """Frictionless sphere on a plane: |x| K d(|x|)^2 / (1 - d(|x|)) = g, independent of the mass."""
    from scipy.optimize import brentq
    xml = """<mujoco><compiler angle="radian"/><option timestep="0.002"/><worldbody>
      <geom name="floor" type="plane" size="5 5 0.1" condim="1"/>
      <body pos="0 0 0.1"><freejoint/><geom type="sphere" size="0.1" condim="1" density="500"/></body>
    </worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m)
    for _ in range(1500):
        od.step()
    od.forward()
    assert od.int("ncon") == 1 and od.int("nefc") == 1
    dist = od.field("con_dist")[0]
    assert np.abs(od.field("qvel")).max() < 1e-7
    d0, dmax, width, mid, power = 0.9, 0.95, 0.001, 0.5, 2.0
    tc, dr = 0.02, 1.0
    K = 1 / (dmax ** 2 * tc ** 2 * dr ** 2)
    def imp(x):
        r = min(x / width, 1.0)
        y = r ** power / mid ** (power - 1) if r <= mid else 1 - (1 - r) ** power / (1 - mid) ** (power - 1)
        return d0 + y * (dmax - d0)
    root = brentq(lambda x: x * K * imp(x) ** 2 / (1 - imp(x)) - 9.81, 1e-9, 0.1)
    assert abs(-dist - root) < 1e-7
    # contact normal force equals the weight
    assert abs(od.field("efc_force")[0] - m.body_mass[1] * 9.81) < 1e-6


def test_joint_limit_row_geometry():
    m = ox.Model.from_xml_string(ox.models.CHEETAH)
    od = OracleData(m)
    j = ox.Model.object_id(m, ox.obj.Joint, "bthigh").index
    qa, da = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j])
    od.field("qpos")[2] = 0.0
    od.field("qpos")[1] = 1.0  # lift off the floor: only the limit is active
    od.field("qpos")[qa] = m.jnt_range[2 * j + 1] + 0.05   # beyond the upper limit
    od.field("qvel")[da] = 50.0                             # and still moving into it (the joint spring alone would pull it back)
    od.forward()
    assert od.int("ncon") == 0 and od.int("nefc") == 1
    J = od.field("efc_J")[:m.nv]
    assert J[da] == -1 and np.count_nonzero(J) == 1
    assert abs(od.field("efc_pos")[0] + 0.05) < 1e-12
    assert od.field("efc_force")[0] > 0 and od.field("qfrc_constraint")[da] < 0


# ---------------------------------------------------------------- D.8 integrators
def _acrobot_energy(m, od):
    od.forward()
    M = dense_M(m, od.field("qM"))
    v = od.field("qvel")
    pe = sum(m.body_mass[i] * 9.81 * od.field("xipos")[3 * i + 2] for i in range(1, m.nbody))
    return 0.5 * v @ M @ v + pe


def test_rk4_is_fourth_order_and_euler_first_order():
    """Trajectory error against a fine-step reference: halving h divides it by ~16 (RK4) and ~2 (Euler)."""
    def run(integrator, h, T=1.0):
        xml = ox.models.ACROBOT.replace('timestep="0.01" integrator="RK4"', f'timestep="{h}" integrator="{integrator}"')
        m = ox.Model.from_xml_string(xml)
        od = OracleData(m)
        od.field("qpos")[:] = [1.0, 0.5]
        e0 = _acrobot_energy(m, od)
        for _ in range(int(round(T / h))):
            od.step()
        return od.field("qpos").copy(), abs(_acrobot_energy(m, od) - e0)
    ref, _ = run("RK4", 0.0005)
    err = {h: np.abs(run("RK4", h)[0] - ref).max() for h in (0.02, 0.01, 0.005)}
    assert 12 < err[0.02] / err[0.01] < 20 and 12 < err[0.01] / err[0.005] < 20
    e1 = {h: np.abs(run("Euler", h)[0] - ref).max() for h in (0.002, 0.001)}
    assert 1.8 < e1[0.002] / e1[0.001] < 2.2
    assert run("RK4", 0.01)[1] < 1e-6 and run("Euler", 0.01)[1] > 1e-2   # energy drift over 1 s


# ---------------------------------------------------------------- D.6 collision known answers
def _one_geom_model(geom, pos, extra=""):
    return f"""<mujoco><compiler angle="radian"/><worldbody>
      <geom name="floor" type="plane" size="5 5 0.1"/>
      <body pos="{pos}"><freejoint/><geom name="g" {geom}/></body>{extra}
    </worldbody></mujoco>"""


def test_collision_known_answers():
    # sphere r=0.1 centred 0.08 above the plane: dist -0.02, pos on the mid-surface, normal +z
    m = ox.Model.from_xml_string(_one_geom_model('type="sphere" size="0.1"', "0.3 0.2 0.08"))
    od = OracleData(m); od.forward()
    assert od.int("ncon") == 1
    assert abs(od.field("con_dist")[0] + 0.02) < 1e-15
    assert np.allclose(od.field("con_pos")[:3], [0.3, 0.2, -0.01])
    fr = od.field("con_frame")[:9].reshape(3, 3)
    assert np.allclose(fr[0], [0, 0, 1]) and np.allclose(fr[1], [0, 1, 0]) and np.allclose(fr[2], [-1, 0, 0])
    # horizontal capsule: two contacts at the end-cap centres, tangent along the capsule axis
    m = ox.Model.from_xml_string(_one_geom_model('type="capsule" fromto="-0.2 0 0 0.2 0 0" size="0.05"', "0 0 0.04"))
    od = OracleData(m); od.forward()
    assert od.int("ncon") == 2
    assert np.allclose(od.field("con_dist")[:2], [-0.01, -0.01])
    xs = sorted([od.field("con_pos")[0], od.field("con_pos")[3]])
    assert np.allclose(xs, [-0.2, 0.2])
    assert abs(abs(od.field("con_frame")[3]) - 1) < 1e-12
    # margin: a sphere hovering inside the margin creates a contact with positive distance, no force
    m = ox.Model.from_xml_string(_one_geom_model('type="sphere" size="0.1" margin="0.05" gap="0.05"', "0 0 0.12"))
    od = OracleData(m); od.forward()
    assert od.int("ncon") == 1 and od.int("nefc") == 0 and abs(od.field("con_dist")[0] - 0.02) < 1e-15
    # box: four bottom corners
    m = ox.Model.from_xml_string(_one_geom_model('type="box" size="0.1 0.2 0.3"', "0 0 0.29"))
    od = OracleData(m); od.forward()
    assert od.int("ncon") == 4 and np.allclose(od.field("con_dist")[:4], -0.01)
    # capsule-capsule, perpendicular, geom1 -> geom2 normal
    xml = """<mujoco><compiler angle="radian"/><worldbody>
      <body pos="0 0 1"><freejoint/><geom type="capsule" fromto="-0.3 0 0 0.3 0 0" size="0.05"/></body>
      <body pos="0 0 1.08"><freejoint/><geom type="capsule" fromto="0 -0.3 0 0 0.3 0" size="0.05"/></body>
    </worldbody></mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m); od.forward()
    assert od.int("ncon") == 1 and abs(od.field("con_dist")[0] + 0.02) < 1e-12
    assert np.allclose(od.field("con_frame")[:3], [0, 0, 1]) and np.allclose(od.field("con_pos")[:3], [0, 0, 1.04])
    # parallel capsules: two contacts
    xml = xml.replace('fromto="0 -0.3 0 0 0.3 0"', 'fromto="-0.2 0 0 0.2 0 0"')
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m); od.forward()
    assert od.int("ncon") == 2 and np.allclose(od.field("con_dist")[:2], -0.02)


def test_accelerometer_closed_form_pendulum():
    """Accelerometer (mj_objectAcceleration in the site frame) on a swinging pendulum: R^T (alpha x r + w x (w x r) - g),
    and -g at rest. Pins the oracle's body-acceleration pass; the kernels are checked against the oracle on the zoo models."""
    xml = """
    <mujoco><compiler angle="radian"/><option timestep="0.002" gravity="0 0 -9.81"/>
      <worldbody><body name="p" pos="0 0 1">
        <joint name="h" type="hinge" axis="0 1 0" damping="0.05"/>
        <geom type="capsule" fromto="0 0 0 0 0 -0.5" size="0.03" density="900"/>
        <site name="imu" pos="0.02 0 -0.4" euler="0 0.3 0"/>
      </body></worldbody>
      <actuator><motor joint="h" gear="1"/></actuator>
      <sensor><accelerometer site="imu"/><gyro site="imu"/></sensor>
    </mujoco>"""
    m = ox.Model.from_xml_string(xml)
    od = OracleData(m)
    for theta, w, u in [(0.0, 0.0, 0.0), (0.7, -1.3, 0.4), (-2.1, 3.0, -1.0)]:
        od.reset()
        od.field("qpos")[0] = theta; od.field("qvel")[0] = w; od.field("ctrl")[0] = u
        od.forward()
        alpha = od.field("qacc")[0]
        c, s = np.cos(theta), np.sin(theta)
        Rb = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])                      # body orientation (rotation about y)
        ce, se = np.cos(0.3), np.sin(0.3)
        Rs = Rb @ np.array([[ce, 0, se], [0, 1, 0], [-se, 0, ce]])             # site orientation
        r = Rb @ np.array([0.02, 0, -0.4])
        wv, av = np.array([0, w, 0.0]), np.array([0, alpha, 0.0])
        a = np.cross(av, r) + np.cross(wv, np.cross(wv, r)) - np.array([0, 0, -9.81])
        assert np.allclose(od.field("sensordata")[:3], Rs.T @ a, rtol=0, atol=1e-10)
        assert np.allclose(od.field("sensordata")[3:6], Rs.T @ wv, atol=1e-12)
