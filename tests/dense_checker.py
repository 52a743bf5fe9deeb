"""Dense whole-step checker: a SECOND, structurally different restatement of mj_step used to pin the oracle.

The oracle (oracle/ox_oracle.cpp) and the CUDA kernels (csrc/ox_stages.cuh) follow MuJoCo's own recursive algorithms
(CRB, RNE, sparse L'DL, primal Newton with exact line search) and share the product's MJCF compiler, so their agreement
proves a faithful port of each other. This module shares none of that machinery:

  * kinematics is a plain product of homogeneous transforms written from the model tables, differentiated by torch
    forward-mode autodiff (exact Jacobians and exact second directional derivatives - no finite differences);
  * the mass matrix is the dense sum  M = sum_b m_b Jc_b' Jc_b + Jw_b' I_b Jw_b (+ armature) and the bias force is the
    projected Newton-Euler (Kane) sum  c = sum_b Jc_b' m_b (a_b - g) + Jw_b' (I_b alpha_b + w_b x I_b w_b), where a_b and
    alpha_b are the accelerations at zero generalised acceleration - no composite bodies, no com frame, no recursion;
  * collision is brute-force geometry (exact segment-segment closest points over the feasible square) on the world-frame
    geoms, with contact parameters re-mixed here from the geom tables (so the compiler's pair tables are checked too);
  * constraint Jacobians are d(point)/dq of the same autodiff kinematics; impedance / reference / regularisation are
    written from the formulas of the MuJoCo documentation as restated in SURVEY.md Appendix A.5-A.6;
  * the constrained acceleration is the minimiser of the convex objective found by a dense damped Newton iteration with a
    backtracking line search (scipy-free, LAPACK solves), run to a gradient norm of ~1e-13;
  * Euler (with the implicit joint-damping solve on the dense M + hB) and RK4 advance the state.

What it shares with the oracle: the compiled model tables as input (body frames, inertias, invweight0 - themselves pinned
by tests/test_compiler.py and the autograd-Lagrangian tests) and the reading of MuJoCo's documentation by the same author.
It is test infrastructure: nothing in the product imports it. tools/make_golden.py runs it to produce tests/golden/*.json.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.func import jacfwd, jvp

torch.set_default_dtype(torch.float64)

MINVAL, MINIMP, MAXIMP = 1e-15, 1e-4, 0.9999
PLANE, SPHERE, CAPSULE, BOX = 0, 2, 3, 6
FREE, BALL, SLIDE, HINGE = 0, 1, 2, 3
DSBL = dict(constraint=1 << 0, limit=1 << 3, contact=1 << 4, passive=1 << 5, gravity=1 << 6, clampctrl=1 << 7,
            filterparent=1 << 9, equality=1 << 1, warmstart=1 << 8, frictionloss=1 << 2, actuation=1 << 10, refsafe=1 << 11, eulerdamp=1 << 13)


def _T(a):
    return torch.as_tensor(np.asarray(a, dtype=np.float64))


def quat2mat_t(q):
    w, x, y, z = q[0], q[1], q[2], q[3]
    return torch.stack([torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
                        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
                        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])])


def hat_t(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def vee(W):
    return np.array([W[2, 1], W[0, 2], W[1, 0]])


def quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat2mat(q):
    return quat2mat_t(_T(q)).numpy()


def mat2quat_pos(R):
    """unit quaternion (w, x, y, z) of a rotation matrix, w >= 0 branch; the largest diagonal term picks the pivot"""
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = 2 * np.sqrt(tr + 1.0)
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = 2 * np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2])
        q = np.array([(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s])
    elif R[1, 1] > R[2, 2]:
        s = 2 * np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2])
        q = np.array([(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s])
    else:
        s = 2 * np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1])
        q = np.array([(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s])
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


class DenseModel:
    """numpy copies of the compiled tables (read through the Python Model mirror)."""
    INT = ["body_parentid", "body_jntadr", "body_jntnum", "jnt_type", "jnt_qposadr", "jnt_dofadr", "jnt_bodyid", "jnt_limited",
           "dof_bodyid", "geom_type", "geom_bodyid", "geom_condim", "geom_priority", "pair_geom1", "pair_geom2", "pair_dim",
           "actuator_trnid", "actuator_trntype", "actuator_gaintype", "actuator_biastype", "actuator_ctrllimited", "actuator_forcelimited",
           "actuator_dyntype", "actuator_actadr", "actuator_actlimited", "body_mocapid", "eq_type", "eq_obj1id", "eq_obj2id",
           "sensor_type", "sensor_objid", "sensor_adr", "site_bodyid", "tendon_adr", "tendon_num", "tendon_limited", "tendon_type", "wrap_objid"]
    REAL = ["qpos0", "qpos_spring", "body_pos", "body_quat", "body_ipos", "body_iquat", "body_mass", "body_inertia", "body_invweight0",
            "jnt_pos", "jnt_axis", "jnt_stiffness", "jnt_range", "jnt_margin", "jnt_solref", "jnt_solimp", "dof_armature", "dof_damping",
            "dof_invweight0", "geom_size", "geom_pos", "geom_quat", "geom_friction", "geom_solmix", "geom_solref", "geom_solimp",
            "geom_margin", "geom_gap", "pair_friction", "pair_solref", "pair_solimp", "pair_margin", "pair_gap", "actuator_gear",
            "actuator_gainprm", "actuator_biasprm", "actuator_ctrlrange", "actuator_forcerange", "actuator_dynprm", "actuator_actrange",
            "eq_solref", "eq_solimp", "eq_data", "site_pos", "site_quat", "wrap_prm", "tendon_range", "tendon_margin", "tendon_solref_lim",
            "tendon_solimp_lim", "tendon_stiffness", "tendon_damping", "tendon_lengthspring", "tendon_invweight0",
            "dof_frictionloss", "dof_solref_fri", "dof_solimp_fri"]

    def __init__(self, model):
        self.m = model
        for k in ("nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "npair", "integrator", "disableflags", "cone", "nmocap", "neq", "nsensor", "nsensordata", "ntendon"):
            setattr(self, k, int(getattr(model, k)))
        for k in ("timestep", "impratio", "density", "viscosity"):
            setattr(self, k, float(getattr(model, k)))
        self.gravity = np.asarray(model.gravity, dtype=np.float64)
        self.wind = np.asarray(model.wind, dtype=np.float64)
        for k in self.INT:
            setattr(self, k, np.asarray(getattr(model, k), dtype=np.int64))
        for k in self.REAL:
            setattr(self, k, np.asarray(getattr(model, k), dtype=np.float64))

    def dis(self, name):
        return bool(self.disableflags & DSBL[name])


# ------------------------------------------------------------------------------------------------ kinematics (autodiff)
def fk(dm: DenseModel, qpos, x, mocap=None):
    """Body frames from generalised coordinates. mocap = (mocap_pos, mocap_quat) for models with mocap bodies. x[nv]: hinge/slide coordinates; for free joints global position + a LOCAL
    rotation vector on top of the quaternion stored in qpos (evaluated at 0); for ball joints a local rotation vector.
    d/dt x = qvel in MuJoCo's convention. Returns P[nbody,3], R[nbody,3,3]."""
    P, R = [torch.zeros(3)], [torch.eye(3)]
    for b in range(1, dm.nbody):
        p = int(dm.body_parentid[b])
        Pb = P[p] + R[p] @ _T(dm.body_pos[3 * b:3 * b + 3])
        Rb = R[p] @ quat2mat_t(_T(dm.body_quat[4 * b:4 * b + 4]))
        if dm.nmocap and dm.body_mocapid[b] >= 0:                        # pose prescribed by the user, not by the tree
            k = int(dm.body_mocapid[b])
            mq = _T(mocap[1][4 * k:4 * k + 4])
            Pb, Rb = _T(mocap[0][3 * k:3 * k + 3]), quat2mat_t(mq / torch.linalg.norm(mq))
        for j in range(int(dm.body_jntadr[b]), int(dm.body_jntadr[b] + dm.body_jntnum[b])):
            jt, qa, da = int(dm.jnt_type[j]), int(dm.jnt_qposadr[j]), int(dm.jnt_dofadr[j])
            ax, jp = _T(dm.jnt_axis[3 * j:3 * j + 3]), _T(dm.jnt_pos[3 * j:3 * j + 3])
            if jt == SLIDE:
                Pb = Pb + Rb @ ax * (x[da] - dm.qpos0[qa])
            elif jt == HINGE:
                anchor = Pb + Rb @ jp
                K = hat_t(ax)
                ang = x[da] - dm.qpos0[qa]
                Rb = Rb @ (torch.eye(3) + torch.sin(ang) * K + (1 - torch.cos(ang)) * (K @ K))
                Pb = anchor - Rb @ jp
            else:
                if jt == FREE:
                    Pb = x[da:da + 3]
                    th, q = x[da + 3:da + 6], _T(qpos[qa + 3:qa + 7])
                    anchor = None
                else:
                    th, q = x[da:da + 3], _T(qpos[qa:qa + 4])
                    anchor = Pb + Rb @ jp
                q = q / torch.linalg.norm(q)
                Hh = hat_t(th)
                local = quat2mat_t(q) @ (torch.eye(3) + Hh + 0.5 * Hh @ Hh)     # exp(hat th) to second order: exact derivatives at th = 0
                Rb = local if jt == FREE else Rb @ local
                if anchor is not None:
                    Pb = anchor - Rb @ jp
        P.append(Pb)
        R.append(Rb)
    return torch.stack(P), torch.stack(R)


def coords(dm: DenseModel, qpos):
    x = np.zeros(dm.nv)
    for j in range(dm.njnt):
        jt, qa, da = int(dm.jnt_type[j]), int(dm.jnt_qposadr[j]), int(dm.jnt_dofadr[j])
        if jt in (SLIDE, HINGE):
            x[da] = qpos[qa]
        elif jt == FREE:
            x[da:da + 3] = qpos[qa:qa + 3]
    return x


class Kin:
    """Frames, exact Jacobians and zero-acceleration accelerations of every body at (qpos, qvel)."""

    def __init__(self, dm: DenseModel, qpos, qvel, mocap=None):
        self.dm = dm
        x, v = _T(coords(dm, qpos)), _T(qvel)
        f = lambda y: fk(dm, qpos, y, mocap)
        (P, R) = f(x)
        dP, dR = jacfwd(f)(x)                                           # [nb,3,nv], [nb,3,3,nv]
        first = lambda y: jvp(f, (y,), (v,))[1]                         # d/dt of (P, R) along qvel
        (vP, vR), (aP, aR) = jvp(first, (x,), (v,))                     # and its derivative along qvel again (qacc = 0)
        self.P, self.R = P.numpy(), R.numpy()
        self.JP = dP.numpy()                                            # translational Jacobian of the body origin
        nb, nv = dm.nbody, dm.nv
        self.JW = np.zeros((nb, 3, nv))                                 # angular Jacobian (world frame)
        for b in range(nb):
            for k in range(nv):
                self.JW[b, :, k] = vee(dR[b, :, :, k].numpy() @ self.R[b].T)
        self.vP, self.aP = vP.numpy(), aP.numpy()
        self.w = np.stack([vee(vR[b].numpy() @ self.R[b].T) for b in range(nb)])
        self.alpha = np.stack([vee(aR[b].numpy() @ self.R[b].T + vR[b].numpy() @ vR[b].numpy().T) for b in range(nb)])
        self.vR, self.aR = vR.numpy(), aR.numpy()

    def point_jac(self, b, p):
        """3 x nv Jacobian of the world position of the material point of body b currently at p."""
        r = p - self.P[b]
        return self.JP[b] - np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]]) @ self.JW[b]


def mass_matrix_and_bias(dm: DenseModel, kin: Kin):
    nv = dm.nv
    M = np.diag(dm.dof_armature.copy())
    c = np.zeros(nv)
    g = np.zeros(3) if dm.dis("gravity") else dm.gravity
    for b in range(1, dm.nbody):
        mb = dm.body_mass[b]
        ipos = dm.body_ipos[3 * b:3 * b + 3]
        Ri = kin.R[b] @ quat2mat(dm.body_iquat[4 * b:4 * b + 4])
        Iw = Ri @ np.diag(dm.body_inertia[3 * b:3 * b + 3]) @ Ri.T
        com = kin.P[b] + kin.R[b] @ ipos
        Jc = kin.point_jac(b, com)
        Jw = kin.JW[b]
        M += mb * Jc.T @ Jc + Jw.T @ Iw @ Jw
        a_com = kin.aP[b] + kin.aR[b] @ ipos                            # second time derivative of the com at qacc = 0
        w = kin.w[b]
        c += Jc.T @ (mb * (a_com - g)) + Jw.T @ (Iw @ kin.alpha[b] + np.cross(w, Iw @ w))
    return M, c


# ------------------------------------------------------------------------------------------------ smooth forces
def sub_quat(qa, qb):
    """3-vector log of qb^-1 * qa (mju_subQuat)."""
    d = quat_mul(np.array([qb[0], -qb[1], -qb[2], -qb[3]]), qa)
    s = np.linalg.norm(d[1:])
    if s < MINVAL:
        return np.zeros(3)
    ang = 2 * np.arctan2(s, d[0])
    if ang > np.pi:
        ang -= 2 * np.pi
    return d[1:] / s * ang


def fluid_force(dm, kin):
    """Drag of the medium (MuJoCo's inertia-box model, computation documentation 'Passive forces'): each body is the box of equal
    mass and principal inertia, moving with the velocity of its centre of mass relative to the wind; Stokes drag of the equivalent
    sphere plus quadratic drag face by face, in the inertial frame; mapped to the joints by the exact (autodiff) Jacobians."""
    f = np.zeros(dm.nv)
    rho, mu = dm.density, dm.viscosity
    if rho <= 0 and mu <= 0:
        return f
    for b in range(1, dm.nbody):
        mass, I = dm.body_mass[b], dm.body_inertia[3 * b:3 * b + 3]
        if mass < MINVAL:
            continue
        ipos = dm.body_ipos[3 * b:3 * b + 3]
        Ri = kin.R[b] @ quat2mat(dm.body_iquat[4 * b:4 * b + 4])
        com = kin.P[b] + kin.R[b] @ ipos
        v = Ri.T @ (kin.vP[b] + kin.vR[b] @ ipos - dm.wind)               # d/dt of the com, in the inertial frame
        w = Ri.T @ kin.w[b]
        box = np.sqrt(np.maximum(MINVAL, I.sum() - 2 * I) / mass * 6.0)  # I_y + I_z - I_x = m/6 * (2 a)^2 / ... -> full side length
        frc, trq = np.zeros(3), np.zeros(3)
        if mu > 0:
            d = box.mean()
            trq -= np.pi * d ** 3 * mu * w
            frc -= 3 * np.pi * d * mu * v
        if rho > 0:
            area = np.array([box[1] * box[2], box[0] * box[2], box[0] * box[1]])
            frc -= 0.5 * rho * area * np.abs(v) * v
            quart = np.array([box[1] ** 4 + box[2] ** 4, box[0] ** 4 + box[2] ** 4, box[0] ** 4 + box[1] ** 4])
            trq -= rho * box * quart * np.abs(w) * w / 64.0
        f += kin.point_jac(b, com).T @ (Ri @ frc) + kin.JW[b].T @ (Ri @ trq)
    return f


def passive_force(dm, qpos, qvel):
    f = np.zeros(dm.nv)
    if dm.dis("passive"):
        return f
    for j in range(dm.njnt):
        k, jt, qa, da = dm.jnt_stiffness[j], int(dm.jnt_type[j]), int(dm.jnt_qposadr[j]), int(dm.jnt_dofadr[j])
        if k == 0:
            continue
        if jt == FREE:
            f[da:da + 3] -= k * (qpos[qa:qa + 3] - dm.qpos_spring[qa:qa + 3])
            qa, da = qa + 3, da + 3
        if jt in (FREE, BALL):
            q = qpos[qa:qa + 4] / np.linalg.norm(qpos[qa:qa + 4])
            f[da:da + 3] -= k * sub_quat(q, dm.qpos_spring[qa:qa + 4])
        else:
            f[da] -= k * (qpos[qa] - dm.qpos_spring[qa])
    f = f - dm.dof_damping * qvel
    for i in range(dm.ntendon):                                          # tendon spring (dead band) and damper through the coefficient vector
        Jt, L = dm._tenJ[i], dm._tenL[i]
        lo, hi = dm.tendon_lengthspring[2 * i:2 * i + 2]
        frc = dm.tendon_stiffness[i] * ((hi - L) if L > hi else (lo - L) if L < lo else 0.0) - dm.tendon_damping[i] * (Jt @ qvel)
        f = f + Jt * frc
    return f


def tendon_jacobian(dm, i):
    return dm._tenJ[i]


def tendons(dm, kin, qpos):
    """lengths and Jacobian rows of all tendons -> dm._tenL, dm._tenJ. Fixed: the coefficient vector; spatial: the path through the
    sites, differentiated through the autodiff point Jacobians of the bodies that carry them."""
    dm._tenJ, dm._tenL = np.zeros((dm.ntendon, dm.nv)), np.zeros(dm.ntendon)
    x = tendon_coords(dm, qpos)
    for i in range(dm.ntendon):
        ws = range(int(dm.tendon_adr[i]), int(dm.tendon_adr[i] + dm.tendon_num[i]))
        if int(dm.tendon_type[i]) == 0:
            for w in ws:
                dm._tenJ[i, int(dm.jnt_dofadr[int(dm.wrap_objid[w])])] += dm.wrap_prm[w]
            dm._tenL[i] = dm._tenJ[i] @ x
        else:
            pts = []
            for w in ws:
                sid = int(dm.wrap_objid[w]); bd = int(dm.site_bodyid[sid])
                pts.append((bd, kin.P[bd] + kin.R[bd] @ dm.site_pos[3 * sid:3 * sid + 3]))
            for (ba, pa), (bb, pb) in zip(pts[:-1], pts[1:]):
                seg = pb - pa
                ln = np.linalg.norm(seg)
                dm._tenL[i] += ln
                if ln > MINVAL:
                    dm._tenJ[i] += seg / ln @ (kin.point_jac(bb, pb) - kin.point_jac(ba, pa))


def tendon_coords(dm, qpos):
    """joint coordinates laid out by dof address (scalar joints only carry a tendon)"""
    x = np.zeros(dm.nv)
    for j in range(dm.njnt):
        if int(dm.jnt_type[j]) in (SLIDE, HINGE):
            x[int(dm.jnt_dofadr[j])] = qpos[int(dm.jnt_qposadr[j])]
    return x


def actuator_force(dm, qpos, qvel, ctrl, act=None):
    """qfrc_actuator, actuator_force, act_dot and the diagonal of d qfrc_actuator / d qvel (used by implicitfast)."""
    f, frc = np.zeros(dm.nv), np.zeros(dm.nu)
    act_dot, dfdv = np.zeros(dm.na), np.zeros(dm.nv)
    if dm.dis("actuation"):
        return f, frc, act_dot, dfdv
    for i in range(dm.nu):
        j, gear = int(dm.actuator_trnid[i]), dm.actuator_gear[i]
        if int(dm.actuator_trntype[i]) == 3:                     # tendon transmission: moment arm = gear * (tendon coefficient vector)
            moment = gear * dm._tenJ[j]
            da = None
            tlen = gear * dm._tenL[j]
        else:
            moment = np.zeros(dm.nv)
            da = int(dm.jnt_dofadr[j])
            moment[da] = gear
        length, vel = (tlen if da is None else moment @ tendon_coords(dm, qpos)), moment @ qvel
        u = ctrl[i]
        if dm.actuator_ctrllimited[i] and not dm.dis("clampctrl"):
            u = min(max(u, dm.actuator_ctrlrange[2 * i]), dm.actuator_ctrlrange[2 * i + 1])
        dyn = int(dm.actuator_dyntype[i])
        if dyn != 0:                                            # stateful: ctrl drives act_dot, act drives the force
            aa = int(dm.actuator_actadr[i])
            act_dot[aa] = u if dyn == 1 else (u - act[aa]) / max(MINVAL, dm.actuator_dynprm[3 * i])
            u = act[aa]
        gp, bp = dm.actuator_gainprm[3 * i:3 * i + 3], dm.actuator_biasprm[3 * i:3 * i + 3]
        gaff, baff = dm.actuator_gaintype[i] == 1, dm.actuator_biastype[i] == 1
        gain = gp[0] + (gp[1] * length + gp[2] * vel if gaff else 0.0)
        bias = bp[0] + bp[1] * length + bp[2] * vel if baff else 0.0
        force = gain * u + bias
        clamped = False
        if dm.actuator_forcelimited[i]:
            lo, hi = dm.actuator_forcerange[2 * i], dm.actuator_forcerange[2 * i + 1]
            clamped = force <= lo or force >= hi
            force = min(max(force, lo), hi)
        frc[i] = force
        f += moment * force
        if not clamped and da is not None:
            dfdv[da] += gear * gear * ((gp[2] * u if gaff else 0.0) + (bp[2] if baff else 0.0))
    return f, frc, act_dot, dfdv


def next_activation(dm, act, act_dot):
    out = np.array(act, float)
    h = dm.timestep
    for i in range(dm.nu):
        dyn = int(dm.actuator_dyntype[i])
        if dyn == 0:
            continue
        aa = int(dm.actuator_actadr[i])
        if dyn == 3:
            tau = max(MINVAL, dm.actuator_dynprm[3 * i])
            out[aa] += act_dot[aa] * tau * (1 - np.exp(-h / tau))
        else:
            out[aa] += h * act_dot[aa]
        if dm.actuator_actlimited[i]:
            out[aa] = min(max(out[aa], dm.actuator_actrange[2 * i]), dm.actuator_actrange[2 * i + 1])
    return out


# ------------------------------------------------------------------------------------------------ collision
def make_frame(n, hint=None):
    n = n / np.linalg.norm(n)
    if hint is None or np.linalg.norm(hint) < 0.5:
        hint = np.array([0.0, 1.0, 0.0]) if -0.5 < n[1] < 0.5 else np.array([0.0, 0.0, 1.0])
    t = hint - n * (n @ hint)
    t = t / np.linalg.norm(t)
    return np.stack([n, t, np.cross(n, t)])


def sphere_sphere(p1, r1, p2, r2, margin):
    d = p2 - p1
    dist = np.linalg.norm(d)
    if dist > margin + r1 + r2:
        return None
    n = d / dist
    return dist - r1 - r2, p1 + n * (r1 + (dist - r1 - r2) / 2), n


def plane_sphere(p0, n, c, r, margin):
    h = n @ (c - p0)
    if h > margin + r:
        return None
    dist = h - r
    return dist, c - n * (r + dist / 2), n


def segment_closest(p1, a1, p2, a2):
    """argmin over s,t in [-1,1] of |p1 + s a1 - p2 - t a2|: the interior stationary point if feasible, else the best of
    the four edges of the square (each a clamped 1-D problem). Brute force on purpose."""
    d = p1 - p2
    A, B, Cc = a1 @ a1, a1 @ a2, a2 @ a2
    best = None

    def consider(s, t):
        nonlocal best
        v = d + s * a1 - t * a2
        val = v @ v
        if best is None or val < best[0] - 1e-300:
            best = (val, s, t)
    det = A * Cc - B * B
    assert abs(det) > 1e-12 * A * Cc, "parallel capsules: outside the dense checker's scope"
    s = (B * (a2 @ d) - Cc * (a1 @ d)) / det
    t = (A * (a2 @ d) - B * (a1 @ d)) / det
    if -1 <= s <= 1 and -1 <= t <= 1:
        consider(s, t)
    else:
        for s0 in (-1.0, 1.0):
            consider(s0, float(np.clip((a2 @ (d + s0 * a1)) / Cc, -1, 1)))
        for t0 in (-1.0, 1.0):
            consider(float(np.clip(-(a1 @ (d - t0 * a2)) / A, -1, 1)), t0)
    return best[1], best[2]


def sphere_box(centre, radius, bpos, bmat, bsize, margin):
    """(dist, pos, normal) or None. World-frame statement: project the centre onto the box by clamping its box coordinates."""
    lc = bmat.T @ (centre - bpos)
    q = np.clip(lc, -bsize, bsize)
    d = np.linalg.norm(q - lc)
    if d - radius > margin:
        return None
    if d <= MINVAL:                                                      # centre inside: leave through the nearest face
        gaps = np.concatenate([lc + bsize, bsize - lc])                  # distances to the -x,-y,-z,+x,+y,+z faces
        order = [0, 3, 1, 4, 2, 5]                                       # the product scans -x,+x,-y,+y,-z,+z and keeps the first minimum
        k = min(order, key=lambda i: (gaps[i], order.index(i)))
        n = np.zeros(3)
        n[k % 3] = 1.0 if k < 3 else -1.0
        depth = gaps[k]
        return -depth - radius, bpos + bmat @ (lc + n * (radius - depth) / 2), bmat @ n
    n = (q - lc) / d
    return d - radius, bpos + bmat @ (0.5 * (q + lc + n * radius)), bmat @ n


def segment_box_closest(lc, la, bsize):
    """argmin over t in [-1,1] of the distance from lc + t la to the box, in EXACT rational arithmetic (the floats are taken as the
    rationals they are): the squared distance is a quadratic between the breakpoints where a coordinate crosses a face, so every
    piece is minimised in closed form and the pieces compared exactly; the leftmost minimiser wins ties. Near-parallel
    configurations make this minimiser ill-conditioned in floating point, which is exactly where a checker must not be sloppy."""
    from fractions import Fraction as Fr
    c, a, sz = [Fr(float(v)) for v in lc], [Fr(float(v)) for v in la], [Fr(float(v)) for v in bsize]

    def f(t):
        tot = Fr(0)
        for k in range(3):
            p = c[k] + t * a[k]
            e = p - max(-sz[k], min(sz[k], p))
            tot += e * e
        return tot

    ts = {Fr(-1), Fr(1)}
    for k in range(3):
        if a[k] != 0:
            for sg in (-1, 1):
                t = (sg * sz[k] - c[k]) / a[k]
                if -1 < t < 1:
                    ts.add(t)
    ts = sorted(ts)
    inside = [t for t in ts if f(t) == 0]                                 # the axis cuts through the box: the middle of the cut
    if inside:                                                            # (entry and exit are breakpoints or segment ends)
        return float((inside[0] + inside[-1]) / 2)
    best_f, best_t = f(Fr(-1)), Fr(-1)
    for t0, t1 in zip(ts[:-1], ts[1:]):
        tm = (t0 + t1) / 2
        f0, fm, f1 = f(t0), f(tm), f(t1)
        h = t1 - t0
        curv = 4 * (f0 - 2 * fm + f1) / (h * h)                          # exact f'' of this piece (second difference over (h/2)^2)
        slope0 = (f1 - f0) / h - curv * h / 2                            # exact f'(t0+)
        cands = [t0, t1]
        if curv > 0:
            cands.append(max(t0, min(t1, t0 - slope0 / curv)))
        for t in sorted(cands):
            ft = f(t)
            if ft < best_f:
                best_f, best_t = ft, t
    return float(best_t)


def box_box(p1, R1, h1, p2, R2, h2, margin):
    """Contacts [(dist, pos, normal)] of two boxes. Contact generation is a DEFINITION, not a law of physics, so this follows the
    product's construction (separating axes, clipping of the incident face, closest points of the supporting edges) - written
    independently, vectorised differently, and checked on its own by the geometric properties in tests/test_box_contacts.py."""
    d = p2 - p1
    A, B = R1.T, R2.T                                                      # rows = axes
    rad = lambda ax, h, L: float(np.sum(h * np.abs(ax @ L)))
    face = []
    for f in range(6):
        L = A[f] if f < 3 else B[f - 3]
        face.append(abs(d @ L) - rad(A, h1, L) - rad(B, h2, L))
    if max(face) > margin:
        return []
    fidx = int(np.argmax(face))                                            # first maximum
    edge = {}
    for i in range(3):
        for j in range(3):
            L = np.cross(A[i], B[j])
            n = np.linalg.norm(L)
            if n < 1e-6:
                continue
            L = L / n
            edge[(i, j)] = abs(d @ L) - rad(A, h1, L) - rad(B, h2, L)
    if edge and max(edge.values()) > margin:
        return []
    scale = max(h1.max(), h2.max())
    if edge:
        (ei, ej), esep = max(edge.items(), key=lambda kv: (kv[1], -3 * kv[0][0] - kv[0][1]))   # first maximum in (i, j) order
        if esep > face[fidx] + 1e-3 * scale:
            L = np.cross(A[ei], B[ej]); L /= np.linalg.norm(L)
            if L @ d < 0:
                L = -L
            c1 = p1 + sum((1.0 if A[k] @ L > 0 else -1.0) * h1[k] * A[k] for k in range(3) if k != ei)
            c2 = p2 + sum((-1.0 if B[k] @ L > 0 else 1.0) * h2[k] * B[k] for k in range(3) if k != ej)
            u, v, w = A[ei], B[ej], c1 - c2
            b, dd, e = u @ v, u @ w, v @ w
            den = 1 - b * b
            sA = float(np.clip((b * e - dd) / den if den > 1e-12 else 0.0, -h1[ei], h1[ei]))
            tB = float(np.clip((e - b * dd) / den if den > 1e-12 else 0.0, -h2[ej], h2[ej]))
            pA, pB = c1 + sA * u, c2 + tB * v
            dist = (pB - pA) @ L
            return [(dist, 0.5 * (pA + pB), L)] if dist <= margin else []
    refA = fidx < 3
    ra = fidx if refA else fidx - 3
    Rax, Iax, Rp, Ip, Rh, Ih = (A, B, p1, p2, h1, h2) if refA else (B, A, p2, p1, h2, h1)
    n = (1.0 if (Ip - Rp) @ Rax[ra] >= 0 else -1.0) * Rax[ra]
    ia = int(np.argmax(np.abs(Iax @ n)))
    isg = -1.0 if Iax[ia] @ n > 0 else 1.0
    iu, iv = (ia + 1) % 3, (ia + 2) % 3
    poly = [Ip + isg * Ih[ia] * Iax[ia] + su * Ih[iu] * Iax[iu] + sv * Ih[iv] * Iax[iv] for su, sv in ((1, 1), (-1, 1), (-1, -1), (1, -1))]
    for side in range(4):
        ta, ps = (ra + 1 + side // 2) % 3, (-1.0 if side % 2 else 1.0)
        ins = lambda pt: Rh[ta] - ps * ((pt - Rp) @ Rax[ta])
        out = []
        for q in range(len(poly)):
            P, Q = poly[q], poly[(q + 1) % len(poly)]
            dp, dq = ins(P), ins(Q)
            if dp >= 0:
                out.append(P)
            if (dp >= 0) != (dq >= 0):
                out.append(P + dp / (dp - dq) * (Q - P))
        poly = out
        if not poly:
            break
    res = []
    for v in poly[:]:
        depth = (v - Rp) @ n - Rh[ra]
        if depth <= margin and len(res) < 8:
            res.append((depth, v - 0.5 * depth * n, n if refA else -n))
    return res


def mix_params(dm, g1, g2):
    """mj_contactParam."""
    p1, p2 = dm.geom_priority[g1], dm.geom_priority[g2]
    margin = max(dm.geom_margin[g1], dm.geom_margin[g2])
    gap = max(dm.geom_gap[g1], dm.geom_gap[g2])
    if p1 != p2:
        g = g1 if p1 > p2 else g2
        fr, sr, si, dim = dm.geom_friction[3 * g:3 * g + 3], dm.geom_solref[2 * g:2 * g + 2], dm.geom_solimp[5 * g:5 * g + 5], int(dm.geom_condim[g])
    else:
        dim = int(max(dm.geom_condim[g1], dm.geom_condim[g2]))
        fr = np.maximum(dm.geom_friction[3 * g1:3 * g1 + 3], dm.geom_friction[3 * g2:3 * g2 + 3])
        s1, s2 = dm.geom_solmix[g1], dm.geom_solmix[g2]
        if s1 >= MINVAL and s2 >= MINVAL:
            mix = s1 / (s1 + s2)
        elif s1 < MINVAL and s2 < MINVAL:
            mix = 0.5
        else:
            mix = 0.0 if s1 < MINVAL else 1.0
        r1, r2 = dm.geom_solref[2 * g1:2 * g1 + 2], dm.geom_solref[2 * g2:2 * g2 + 2]
        sr = mix * r1 + (1 - mix) * r2 if (r1[0] > 0 and r2[0] > 0) else np.minimum(r1, r2)
        si = mix * dm.geom_solimp[5 * g1:5 * g1 + 5] + (1 - mix) * dm.geom_solimp[5 * g2:5 * g2 + 5]
    fr5 = np.array([fr[0], fr[0], fr[1], fr[2], fr[2]])
    return dict(dim=dim, friction=fr5, solref=np.asarray(sr, float), solimp=np.asarray(si, float), margin=margin, gap=gap)


def collide(dm: DenseModel, kin: Kin):
    """List of contacts dict(dist, pos, frame[3,3], pair) over the model's candidate pair list."""
    out = []
    if dm.dis("contact") or dm.dis("constraint"):
        return out
    gpos = [kin.P[dm.geom_bodyid[g]] + kin.R[dm.geom_bodyid[g]] @ dm.geom_pos[3 * g:3 * g + 3] for g in range(dm.ngeom)]
    gmat = [kin.R[dm.geom_bodyid[g]] @ quat2mat(dm.geom_quat[4 * g:4 * g + 4]) for g in range(dm.ngeom)]
    for p in range(dm.npair):
        g1, g2 = int(dm.pair_geom1[p]), int(dm.pair_geom2[p])
        t1, t2 = int(dm.geom_type[g1]), int(dm.geom_type[g2])
        assert t1 <= t2
        prm = mix_params(dm, g1, g2)
        margin = prm["margin"]
        s1, s2 = dm.geom_size[3 * g1:3 * g1 + 3], dm.geom_size[3 * g2:3 * g2 + 3]
        found = []
        if t1 == PLANE:
            n = gmat[g1][:, 2]
            if t2 == SPHERE:
                r = plane_sphere(gpos[g1], n, gpos[g2], s2[0], margin)
                if r: found.append((r, None))
            elif t2 == CAPSULE:
                ax = gmat[g2][:, 2]
                for sgn in (1.0, -1.0):
                    r = plane_sphere(gpos[g1], n, gpos[g2] + sgn * ax * s2[1], s2[0], margin)
                    if r: found.append((r, ax))
            elif t2 == BOX:
                cnt = 0
                for i in range(8):
                    corner = gmat[g2] @ (np.array([1 if i & 1 else -1, 1 if i & 2 else -1, 1 if i & 4 else -1]) * s2)
                    ld, dist0 = n @ corner, n @ (gpos[g2] - gpos[g1])
                    if dist0 + ld > margin or ld > 0 or cnt >= 4:
                        continue
                    dist = dist0 + ld
                    found.append(((dist, corner + gpos[g2] - n * dist / 2, n), None))
                    cnt += 1
            else:
                raise NotImplementedError((t1, t2))
        elif t1 == SPHERE and t2 == SPHERE:
            r = sphere_sphere(gpos[g1], s1[0], gpos[g2], s2[0], margin)
            if r: found.append((r, None))
        elif t1 == SPHERE and t2 == CAPSULE:
            ax = gmat[g2][:, 2]
            x = float(np.clip(ax @ (gpos[g1] - gpos[g2]), -s2[1], s2[1]))
            r = sphere_sphere(gpos[g1], s1[0], gpos[g2] + ax * x, s2[0], margin)
            if r: found.append((r, None))
        elif t1 == BOX and t2 == BOX:
            for r in box_box(gpos[g1], gmat[g1], s1, gpos[g2], gmat[g2], s2, margin):
                found.append((r, None))
        elif t1 == SPHERE and t2 == BOX:
            r = sphere_box(gpos[g1], s1[0], gpos[g2], gmat[g2], s2, margin)
            if r: found.append((r, None))
        elif t1 == CAPSULE and t2 == BOX:
            ax = gmat[g1][:, 2]
            lc, la = gmat[g2].T @ (gpos[g1] - gpos[g2]), gmat[g2].T @ ax * s1[1]
            tstar = segment_box_closest(lc, la, s2)
            for t in (tstar, 1.0 if tstar <= 0 else -1.0):               # closest point of the axis, then the far end cap
                r = sphere_box(gpos[g1] + t * s1[1] * ax, s1[0], gpos[g2], gmat[g2], s2, margin)
                if r: found.append((r, ax))
        elif t1 == CAPSULE and t2 == CAPSULE:
            a1, a2 = gmat[g1][:, 2] * s1[1], gmat[g2][:, 2] * s2[1]
            s, t = segment_closest(gpos[g1], a1, gpos[g2], a2)
            r = sphere_sphere(gpos[g1] + s * a1, s1[0], gpos[g2] + t * a2, s2[0], margin)
            if r: found.append((r, None))
        else:
            raise NotImplementedError((t1, t2))
        for (dist, pos, n), hint in found:
            out.append(dict(dist=dist, pos=pos, frame=make_frame(n, hint), pair=p, g1=g1, g2=g2, prm=prm))
    return out


# ------------------------------------------------------------------------------------------------ constraints
def impedance(solimp, x):
    d0, dmax, width, mid, power = solimp
    if d0 == dmax or width <= MINVAL:
        d = 0.5 * (d0 + dmax)
    else:
        r = abs(x) / width
        if r >= 1:
            d = dmax
        elif r <= 0:
            d = d0
        else:
            if power == 1:
                y = r
            elif r <= mid:
                y = r ** power / mid ** (power - 1)
            else:
                y = 1 - (1 - r) ** power / (1 - mid) ** (power - 1)
            d = d0 + y * (dmax - d0)
    return min(max(d, MINIMP), MAXIMP)


def row_params(dm, solref, solimp, pos, margin, diag_approx, vel):
    d = impedance(solimp, pos - margin)
    dmax = solimp[1]
    if solref[0] > 0:
        tc, dr = solref
        if not dm.dis("refsafe"):
            tc = max(tc, 2 * dm.timestep)
        K = 1 / max(MINVAL, dmax * dmax * tc * tc * dr * dr)
        Bd = 2 / max(MINVAL, dmax * tc)
    else:
        K, Bd = -solref[0] / max(MINVAL, dmax * dmax), -solref[1] / max(MINVAL, dmax)
    aref = -Bd * vel - K * d * (pos - margin)
    R = max(MINVAL, (1 - d) / d * diag_approx)
    return aref, R


def constraints(dm: DenseModel, kin: Kin, qpos, qvel, contacts, eq_active=None):
    """Rows (J, D, aref) and ne = number of leading equality rows (quadratic on both sides)."""
    J, D, aref = [], [], []
    dm._friction_pairs = []                                              # (row of edge +, row of edge -) of every pyramidal friction direction
    dm._ell = []                                                         # elliptic contacts: (first row, dim, mu, friction[:dim-1])
    cart = dm._row_cart = []                                             # per row: (body+, point+, body-, point-, direction) of the Cartesian force it is, or None
    if dm.dis("constraint"):
        return np.zeros((0, dm.nv)), np.zeros(0), np.zeros(0), 0
    if dm.neq and not dm.dis("equality"):
        for i in range(dm.neq):
            if eq_active is not None and not eq_active[i]:
                continue
            data, sr, si = dm.eq_data[11 * i:11 * i + 11], dm.eq_solref[2 * i:2 * i + 2], dm.eq_solimp[5 * i:5 * i + 5]
            if int(dm.eq_type[i]) in (0, 1):                             # connect / weld: the two anchors coincide
                b1, b2 = int(dm.eq_obj1id[i]), int(dm.eq_obj2id[i])
                p1, p2 = kin.P[b1] + kin.R[b1] @ data[0:3], kin.P[b2] + kin.R[b2] @ data[3:6]
                Jd = kin.point_jac(b1, p1) - kin.point_jac(b2, p2)
                diag = dm.body_invweight0[2 * b1] + dm.body_invweight0[2 * b2]
                for k in range(3):
                    a, R = row_params(dm, sr, si, p1[k] - p2[k], 0.0, diag, Jd[k] @ qvel)
                    J.append(Jd[k]); D.append(1 / R); aref.append(a); cart.append((b1, p1, b2, p2, np.eye(3)[k]))
                if int(dm.eq_type[i]) == 1:                              # weld: orientation error e = R2' R1 Rrel, residual = ts * imag(quat(e))
                    Rrel, ts = quat2mat(data[6:10]), data[10]
                    qerr = lambda R1, R2: mat2quat_pos(R2.T @ R1 @ Rrel)[1:] * ts
                    res = qerr(kin.R[b1], kin.R[b2])
                    # d residual / dt for a relative angular velocity w (world frame): rotate body 1 by a small w dt and difference it
                    # numerically? no - exact: e' = R2' hat(w1 - w2) R1 Rrel; imag(quat)' = 1/2 (E_w + trace-part) ... use the quaternion form
                    q1, q2 = mat2quat_pos(kin.R[b1]), mat2quat_pos(kin.R[b2])
                    quat = quat_mul(q1, mat2quat_pos(Rrel))
                    if quat_mul(np.array([q2[0], -q2[1], -q2[2], -q2[3]]), quat)[0] < 0:
                        quat = -quat                                       # same branch as the residual (w >= 0)
                    G = np.stack([0.5 * ts * quat_mul(quat_mul(np.array([q2[0], -q2[1], -q2[2], -q2[3]]), np.r_[0.0, np.eye(3)[c]]), quat)[1:] for c in range(3)], axis=1)
                    Jr = G @ (kin.JW[b1] - kin.JW[b2])
                    rdiag = dm.body_invweight0[2 * b1 + 1] + dm.body_invweight0[2 * b2 + 1]
                    for k in range(3):
                        a, R = row_params(dm, sr, si, res[k], 0.0, rdiag, Jr[k] @ qvel)
                        J.append(Jr[k]); D.append(1 / R); aref.append(a); cart.append((b1, p1, b2, p2, np.zeros(3), G[k]))
            else:                                                        # joint: q1 follows a quartic polynomial of q2
                j1, j2 = int(dm.eq_obj1id[i]), int(dm.eq_obj2id[i])
                q1, d1 = int(dm.jnt_qposadr[j1]), int(dm.jnt_dofadr[j1])
                row = np.zeros(dm.nv)
                row[d1] = 1.0
                pos, diag = qpos[q1] - dm.qpos0[q1], dm.dof_invweight0[d1]
                if j2 >= 0:
                    q2, d2 = int(dm.jnt_qposadr[j2]), int(dm.jnt_dofadr[j2])
                    dq = qpos[q2] - dm.qpos0[q2]
                    pos -= np.polyval(data[4::-1], dq)
                    row[d2] = -np.polyval(np.polyder(np.poly1d(data[4::-1])).coeffs, dq)
                    diag += dm.dof_invweight0[d2]
                else:
                    pos -= data[0]
                a, R = row_params(dm, sr, si, pos, 0.0, diag, row @ qvel)
                J.append(row); D.append(1 / R); aref.append(a); cart.append(None)
    ne = len(J)
    dm._floss = []                                                       # frictionloss of each dry-friction row (they follow the equality rows)
    if not dm.dis("frictionloss"):
        for i in range(dm.nv):
            if dm.dof_frictionloss[i] > 0:
                row = np.zeros(dm.nv)
                row[i] = 1.0
                a, R = row_params(dm, dm.dof_solref_fri[2 * i:2 * i + 2], dm.dof_solimp_fri[5 * i:5 * i + 5], 0.0, 0.0, dm.dof_invweight0[i], qvel[i])
                J.append(row); D.append(1 / R); aref.append(a); cart.append(None); dm._floss.append(dm.dof_frictionloss[i])
    if not dm.dis("limit"):
        for j in range(dm.njnt):
            if dm.jnt_limited[j] and int(dm.jnt_type[j]) == BALL:            # limit on the rotation angle: J = -(unit rotation axis)
                qa, da = int(dm.jnt_qposadr[j]), int(dm.jnt_dofadr[j])
                rv = sub_quat(qpos[qa:qa + 4] / np.linalg.norm(qpos[qa:qa + 4]), np.array([1.0, 0, 0, 0]))   # log of the joint rotation
                ang = np.linalg.norm(rv)
                dist = max(dm.jnt_range[2 * j], dm.jnt_range[2 * j + 1]) - ang
                if dist < dm.jnt_margin[j]:
                    row = np.zeros(dm.nv)
                    row[da:da + 3] = -rv / ang if ang > MINVAL else 0.0
                    a, R = row_params(dm, dm.jnt_solref[2 * j:2 * j + 2], dm.jnt_solimp[5 * j:5 * j + 5], dist, dm.jnt_margin[j],
                                      dm.dof_invweight0[da], row @ qvel)
                    J.append(row); D.append(1 / R); aref.append(a); cart.append(None)
                continue
            if not dm.jnt_limited[j] or int(dm.jnt_type[j]) not in (SLIDE, HINGE):
                continue
            q, da = qpos[int(dm.jnt_qposadr[j])], int(dm.jnt_dofadr[j])
            for sign, dist in ((1.0, q - dm.jnt_range[2 * j]), (-1.0, dm.jnt_range[2 * j + 1] - q)):
                if dist < dm.jnt_margin[j]:
                    row = np.zeros(dm.nv)
                    row[da] = sign
                    a, R = row_params(dm, dm.jnt_solref[2 * j:2 * j + 2], dm.jnt_solimp[5 * j:5 * j + 5], dist, dm.jnt_margin[j],
                                      dm.dof_invweight0[da], row @ qvel)
                    J.append(row); D.append(1 / R); aref.append(a); cart.append(None)
    if not dm.dis("limit"):
        for i in range(dm.ntendon):
            if not dm.tendon_limited[i]:
                continue
            Jt, L = dm._tenJ[i], dm._tenL[i]
            for sign, dist in ((1.0, L - dm.tendon_range[2 * i]), (-1.0, dm.tendon_range[2 * i + 1] - L)):
                if dist < dm.tendon_margin[i]:
                    row = sign * Jt
                    a, R = row_params(dm, dm.tendon_solref_lim[2 * i:2 * i + 2], dm.tendon_solimp_lim[5 * i:5 * i + 5], dist, dm.tendon_margin[i],
                                      dm.tendon_invweight0[i], row @ qvel)
                    J.append(row); D.append(1 / R); aref.append(a); cart.append(None)
    for c in contacts:
        prm = c["prm"]
        incl = prm["margin"] - prm["gap"]
        if c["dist"] >= incl:
            continue
        b1, b2 = int(dm.geom_bodyid[c["g1"]]), int(dm.geom_bodyid[c["g2"]])
        Jd = c["frame"] @ (kin.point_jac(b2, c["pos"]) - kin.point_jac(b1, c["pos"]))      # rows: normal, tangent 1, tangent 2
        tran = dm.body_invweight0[2 * b1] + dm.body_invweight0[2 * b2]
        if prm["dim"] == 1:
            a, R = row_params(dm, prm["solref"], prm["solimp"], c["dist"], incl, tran, Jd[0] @ qvel)
            J.append(Jd[0]); D.append(1 / R); aref.append(a); cart.append((b2, c["pos"], b1, c["pos"], c["frame"][0]))
        elif dm.cone == 1:                                               # elliptic cone: rows are the contact-frame components themselves
            dim = prm["dim"]
            Jw = c["frame"] @ (kin.JW[b2] - kin.JW[b1])
            comp = [Jd[0], Jd[1], Jd[2], Jw[0], Jw[1], Jw[2]][:dim]
            mu = prm["friction"][0] * np.sqrt(1 / dm.impratio)
            a0_, Rn = row_params(dm, prm["solref"], prm["solimp"], c["dist"], incl, tran, comp[0] @ qvel)
            dm._ell.append((len(J), dim, mu, np.array(prm["friction"][:dim - 1])))
            zero = np.zeros(3)
            for k in range(dim):
                if k == 0:
                    a_k, D_k = a0_, 1 / Rn
                else:                                                    # friction rows: no position term, D_k = D_n f_k^2 / mu^2
                    a_k, _ = row_params(dm, prm["solref"], prm["solimp"], 0.0, 0.0, tran, comp[k] @ qvel)
                    D_k = prm["friction"][k - 1] ** 2 / (Rn * mu * mu)
                lin = c["frame"][k] if k < 3 else zero
                tor = c["frame"][k - 3] if k >= 3 else zero
                J.append(comp[k]); D.append(D_k); aref.append(a_k); cart.append((b2, c["pos"], b1, c["pos"], lin, tor))
        else:
            dim = prm["dim"]
            assert dim in (3, 4, 6)
            # friction directions 1, 2: tangents (relative linear velocity of the contact point); 3: torsion about the normal,
            # 4, 5: rolling about the tangents (relative ANGULAR velocity of the two bodies)
            Jw = c["frame"] @ (kin.JW[b2] - kin.JW[b1])
            Jk = [Jd[1], Jd[2], Jw[0], Jw[1], Jw[2]][:dim - 1]
            zero = np.zeros(3)
            rows, dirs = [], []
            for k in range(dim - 1):
                for s in (1.0, -1.0):
                    rows.append(Jd[0] + s * prm["friction"][k] * Jk[k])
                    lin = c["frame"][0] + (s * prm["friction"][k] * c["frame"][1 + k] if k < 2 else zero)
                    tor = s * prm["friction"][k] * c["frame"][k - 2] if k >= 2 else zero
                    dirs.append((lin, tor))
            _, R0 = row_params(dm, prm["solref"], prm["solimp"], c["dist"], incl, tran + prm["friction"][0] ** 2 * tran, rows[0] @ qvel)
            mu = prm["friction"][0] * np.sqrt(1 / dm.impratio)
            Rpy = 2 * mu * mu * R0
            dm._friction_pairs += [(len(J) + 2 * k, len(J) + 2 * k + 1) for k in range(dim - 1)]
            for r, (lin, tor) in zip(rows, dirs):
                a, _ = row_params(dm, prm["solref"], prm["solimp"], c["dist"], incl, tran, r @ qvel)
                J.append(r); D.append(1 / Rpy); aref.append(a); cart.append((b2, c["pos"], b1, c["pos"], lin, tor))
    if not J:
        return np.zeros((0, dm.nv)), np.zeros(0), np.zeros(0), 0
    return np.array(J), np.array(D), np.array(aref), ne


def solve_qacc_autodiff(M, qfrc_smooth, J, D, aref, ne, floss, ell):
    """Elliptic cones: the objective is written down once (quadratic gauss term + row penalties + the three-zone cone cost of
    every elliptic contact) and torch differentiates it - gradient and Hessian by autograd, damped Newton on top. The product and
    the oracle use hand-derived cone Hessians; nothing of that derivation is shared here."""
    n, nf = J.shape[0], len(floss)
    Mt, Jt, Dt, at_ = _T(M), _T(J), _T(D), _T(aref)
    a0 = np.linalg.solve(M, qfrc_smooth)
    a0t = _T(a0)
    inell = np.zeros(n, bool)
    for r0, dim, _, _ in ell:
        inell[r0:r0 + dim] = True
    two_sided = _T((np.arange(n) < ne + nf).astype(float))
    fl = np.zeros(n)
    fl[ne:ne + nf] = floss
    plain = _T((~inell).astype(float))

    def objective(a):
        x = Jt @ a - at_
        c = 0.5 * (a - a0t) @ Mt @ (a - a0t)
        quad = torch.where(two_sided > 0, x, torch.clamp(x, max=0.0))
        rowc = 0.5 * Dt * quad ** 2
        if nf:
            flt, band = _T(fl), _T(np.where(fl > 0, fl / D, np.inf))
            rowc = torch.where(torch.abs(x) >= band, flt * (torch.abs(x) - 0.5 * flt / Dt), rowc)
        c = c + torch.sum(rowc * plain)
        for r0, dim, mu, fri in ell:
            xc = x[r0:r0 + dim]
            N = xc[0] * mu
            U = xc[1:] * _T(fri)
            Tn = torch.sqrt(torch.sum(U * U))
            if N >= mu * Tn:
                continue
            if mu * N + Tn <= 0:
                c = c + 0.5 * torch.sum(Dt[r0:r0 + dim] * xc ** 2)
            else:
                c = c + 0.5 * Dt[r0] / (mu * mu * (1 + mu * mu)) * (N - mu * Tn) ** 2
        return c

    a = _T(a0.copy())
    for _ in range(200):
        a = a.detach().requires_grad_(True)
        c = objective(a)
        g, = torch.autograd.grad(c, a, create_graph=True)
        if float(torch.linalg.norm(g)) <= 1e-13 * max(1.0, float(np.linalg.norm(M @ a0))):
            break
        H = torch.stack([torch.autograd.grad(g[i], a, retain_graph=True)[0] for i in range(len(a))])
        step = -torch.linalg.solve(H.detach() + 1e-300 * torch.eye(len(a)), g.detach())
        c0, t, gs = float(c), 1.0, float(g.detach() @ step)
        while float(objective((a + t * step).detach())) > c0 + 1e-4 * t * gs and t > 1e-12:
            t *= 0.5
        if t <= 1e-12:
            break
        a = (a + t * step).detach()
    # per-row forces = -d(penalty)/dx at the solution, again by autograd
    aa = a.detach().numpy()
    xv = (J @ aa - aref)
    xt = _T(xv).requires_grad_(True)
    quad = torch.where(two_sided > 0, xt, torch.clamp(xt, max=0.0))
    rowc = 0.5 * Dt * quad ** 2
    if nf:
        flt, band = _T(fl), _T(np.where(fl > 0, fl / D, np.inf))
        rowc = torch.where(torch.abs(xt) >= band, flt * (torch.abs(xt) - 0.5 * flt / Dt), rowc)
    pc = torch.sum(rowc * plain)
    for r0, dim, mu, fri in ell:
        xc = xt[r0:r0 + dim]
        N = xc[0] * mu
        U = xc[1:] * _T(fri)
        Tn = torch.sqrt(torch.sum(U * U))
        if N >= mu * Tn:
            continue
        if mu * N + Tn <= 0:
            pc = pc + 0.5 * torch.sum(Dt[r0:r0 + dim] * xc ** 2)
        else:
            pc = pc + 0.5 * Dt[r0] / (mu * mu * (1 + mu * mu)) * (N - mu * Tn) ** 2
    force = np.zeros(n)
    if pc.requires_grad:
        gx, = torch.autograd.grad(pc, xt, allow_unused=True)
        if gx is not None:
            force = -gx.numpy()
    return aa, force


def solve_qacc(M, qfrc_smooth, J, D, aref, ne=0, floss=()):
    """argmin_a 1/2 (a-a0)' M (a-a0) + sum_r 1/2 D_r min(0, J_r a - aref_r)^2 by damped Newton with backtracking."""
    a0 = np.linalg.solve(M, qfrc_smooth)
    if J.shape[0] == 0:
        return a0, np.zeros(0)
    n, nf = J.shape[0], len(floss)
    two_sided = np.arange(n) < ne + nf                                   # equality rows (and friction rows inside their band): 1/2 D x^2 for either sign
    fl = np.zeros(n)
    fl[ne:ne + nf] = floss
    band = np.where(fl > 0, fl / D, np.inf)                              # dry friction: Huber cost, quadratic for |x| < R floss, slope floss beyond

    def pen_cost(x):
        quad = np.where(two_sided, x, np.minimum(0.0, x))
        lin = np.abs(x) >= band
        return np.sum(np.where(lin, fl * (np.abs(x) - 0.5 * fl / D), 0.5 * D * quad ** 2))

    def pen_force(x):
        lin = np.abs(x) >= band
        return np.where(lin, -np.sign(x) * fl, np.where((x < 0) | two_sided, -D * x, 0.0))

    cost = lambda a: 0.5 * (a - a0) @ M @ (a - a0) + pen_cost(J @ a - aref)
    a = a0.copy()
    for _ in range(500):
        jar = J @ a - aref
        act = ((jar < 0) | two_sided) & (np.abs(jar) < band)
        g = M @ (a - a0) - J.T @ pen_force(jar)
        if np.linalg.norm(g) <= 1e-13 * max(1.0, np.linalg.norm(M @ a0)):
            break
        H = M + (J[act].T * D[act]) @ J[act]
        step = -np.linalg.solve(H, g)
        c0, t = cost(a), 1.0
        while cost(a + t * step) > c0 + 1e-4 * t * (g @ step) and t > 1e-12:
            t *= 0.5
        if t <= 1e-12:
            break
        a = a + t * step
    jar = J @ a - aref
    return a, pen_force(jar)


def solve_dual(dm, M, qfrc_smooth, J, D, aref, ne, warmstart, pgs):
    """MuJoCo's dual solvers written the way mj_solPGS / mj_solNoSlip are: on the EXPLICIT matrix A = J M^-1 J' (the product keeps
    it implicit). pgs=True: projected Gauss-Seidel on AR = A + diag(R); then, if the model asks for it, the noslip pass on the
    friction dimensions with the unregularised A. pgs=False: noslip only, starting from the forces of the primal solution."""
    nv, n = dm.nv, J.shape[0]
    a0 = np.linalg.solve(M, qfrc_smooth)
    if n == 0:
        return a0, np.zeros(0), 0
    A = J @ np.linalg.solve(M, J.T)
    R = 1 / D
    bvec = J @ a0 - aref
    scale = 1 / (float(dm.m.meaninertia) * max(1, nv))
    two_sided = np.arange(n) < ne
    nf = len(dm._floss)
    fl = np.zeros(n)
    fl[ne:ne + nf] = dm._floss
    niter = 0
    if pgs:
        f = np.zeros(n)
        if not dm.dis("warmstart"):
            jar = J @ warmstart - aref
            fw = np.where((jar < 0) | two_sided | (fl > 0), -D * jar, 0.0)
            fw = np.where(fl > 0, np.clip(fw, -fl, fl), fw)
            if 0.5 * fw @ (A @ fw + R * fw) + fw @ bvec < 0:
                f = fw
        AR = A + np.diag(R)
        for niter in range(1, int(dm.m.iterations) + 1):
            improvement = 0.0
            for r in range(n):
                res = AR[r] @ f + bvec[r]
                new = f[r] - res / AR[r, r]
                if fl[r] > 0:
                    new = min(fl[r], max(-fl[r], new))
                elif not two_sided[r]:
                    new = max(0.0, new)
                delta = new - f[r]
                f[r] = new
                improvement -= 0.5 * delta * delta * AR[r, r] + delta * res
            if improvement * scale < float(dm.m.tolerance):
                break
    else:
        f = warmstart                                                   # the primal solver's forces
    for _ in range(int(dm.m.noslip_iterations)):
        improvement = 0.0
        for r in range(ne, ne + nf):                                      # dry-friction rows, unregularised
            if A[r, r] < MINVAL:
                continue
            res = A[r] @ f + bvec[r]
            new = min(fl[r], max(-fl[r], f[r] - res / A[r, r]))
            delta = new - f[r]
            f[r] = new
            improvement -= 0.5 * delta * delta * A[r, r] + delta * res
        for ra, rb in dm._friction_pairs:
            K = A[ra, ra] + A[rb, rb] - 2 * A[ra, rb]
            if K < MINVAL:
                continue
            res = A @ f + bvec
            mid, yold = 0.5 * (f[ra] + f[rb]), 0.5 * (f[ra] - f[rb])
            y = min(mid, max(-mid, yold - (res[ra] - res[rb]) / K))
            delta = y - yold
            f[ra], f[rb] = mid + y, mid - y
            improvement -= 0.5 * delta * delta * K + delta * (res[ra] - res[rb])
        if improvement * scale < float(dm.m.noslip_tolerance):
            break
    return a0 + np.linalg.solve(M, J.T @ f), f, niter


def force_torque_sensors(dm, kin, qacc, force, xfrc_applied):
    """force / torque sensors WITHOUT spatial algebra: the wrench the subtree of the site's body exchanges with the parent is the
    rate of change of the subtree's momentum (Newton-Euler per body, from the autodiff accelerations) minus every external Cartesian
    force on the subtree (gravity, xfrc_applied, contact and connect-constraint forces read off the solver's row forces), taken
    about the site and expressed in the site frame."""
    out = {}
    g = np.zeros(3) if dm.dis("gravity") else dm.gravity
    for s in range(dm.nsensor):
        ty = int(dm.sensor_type[s])
        if ty not in (4, 5):
            continue
        site = int(dm.sensor_objid[s])
        B = int(dm.site_bodyid[site])
        P = kin.P[B] + kin.R[B] @ dm.site_pos[3 * site:3 * site + 3]
        Rs = kin.R[B] @ quat2mat(dm.site_quat[4 * site:4 * site + 4])
        sub = []
        for k in range(1, dm.nbody):
            a = k
            while a > 0 and a != B:
                a = int(dm.body_parentid[a])
            if a == B:
                sub.append(k)
        F, Tq = np.zeros(3), np.zeros(3)
        for k in sub:
            mk, ipos = dm.body_mass[k], dm.body_ipos[3 * k:3 * k + 3]
            Ri = kin.R[k] @ quat2mat(dm.body_iquat[4 * k:4 * k + 4])
            Iw = Ri @ np.diag(dm.body_inertia[3 * k:3 * k + 3]) @ Ri.T
            com = kin.P[k] + kin.R[k] @ ipos
            a_com = kin.point_jac(k, com) @ qacc + kin.aP[k] + kin.aR[k] @ ipos
            alpha = kin.JW[k] @ qacc + kin.alpha[k]
            w = kin.w[k]
            F += mk * (a_com - g)
            Tq += Iw @ alpha + np.cross(w, Iw @ w) + np.cross(com - P, mk * (a_com - g))
            if xfrc_applied is not None:
                xf = xfrc_applied[6 * k:6 * k + 6]
                F -= xf[:3]
                Tq -= xf[3:] + np.cross(com - P, xf[:3])
        for r, info in enumerate(dm._row_cart):
            if info is None or force[r] == 0:
                continue
            bp, pp, bm, pm, dirv = info[:5]
            tor = info[5] if len(info) > 5 else np.zeros(3)                # torsional / rolling friction: a pure torque pair
            if bp in sub:
                F -= force[r] * dirv; Tq -= np.cross(pp - P, force[r] * dirv) + force[r] * tor
            if bm in sub:
                F += force[r] * dirv; Tq += np.cross(pm - P, force[r] * dirv) + force[r] * tor
        out[int(dm.sensor_adr[s])] = Rs.T @ (F if ty == 4 else Tq)
    return out


# ------------------------------------------------------------------------------------------------ whole step
def forward(dm: DenseModel, qpos, qvel, ctrl, qfrc_applied=None, xfrc_applied=None, act=None, mocap=None, eq_active=None, warmstart=None):
    qpos, qvel = np.asarray(qpos, float), np.asarray(qvel, float)
    act = np.zeros(dm.na) if act is None else np.asarray(act, float)
    kin = Kin(dm, qpos, qvel, mocap)
    tendons(dm, kin, qpos)
    M, c = mass_matrix_and_bias(dm, kin)
    fa, frc, act_dot, dfdv = actuator_force(dm, qpos, qvel, np.asarray(ctrl, float), act)
    f = passive_force(dm, qpos, qvel) - c + fa
    if not dm.dis("passive"):
        f = f + fluid_force(dm, kin)
        if dm.m.ngravcomp and not dm.dis("gravity"):                      # gravity compensation: a counter-weight force at each body's com
            gc = np.asarray(dm.m.body_gravcomp, float)
            for b in range(1, dm.nbody):
                if gc[b]:
                    com = kin.P[b] + kin.R[b] @ dm.body_ipos[3 * b:3 * b + 3]
                    f = f + kin.point_jac(b, com).T @ (-dm.gravity * dm.body_mass[b] * gc[b])
    if qfrc_applied is not None:
        f = f + qfrc_applied
    if xfrc_applied is not None:
        for b in range(1, dm.nbody):
            w = xfrc_applied[6 * b:6 * b + 6]
            if np.any(w != 0):
                com = kin.P[b] + kin.R[b] @ dm.body_ipos[3 * b:3 * b + 3]
                f = f + kin.point_jac(b, com).T @ w[:3] + kin.JW[b].T @ w[3:]
    cons = collide(dm, kin)
    J, D, aref, ne = constraints(dm, kin, qpos, qvel, cons, eq_active)
    solver, noslip = int(dm.m.solver), int(dm.m.noslip_iterations)
    if solver == 0:                                                      # PGS (+ noslip)
        qacc, force, _ = solve_dual(dm, M, f, J, D, aref, ne, np.zeros(dm.nv) if warmstart is None else np.asarray(warmstart, float), True)
    else:
        if dm._ell:
            qacc, force = solve_qacc_autodiff(M, f, J, D, aref, ne, dm._floss, dm._ell)
        else:
            qacc, force = solve_qacc(M, f, J, D, aref, ne, dm._floss)
        if noslip and len(force):
            qacc, force, _ = solve_dual(dm, M, f, J, D, aref, ne, force.copy(), False)
    ft = force_torque_sensors(dm, kin, qacc, force, xfrc_applied) if dm.nsensor else {}
    return dict(ft_sensors=ft, qacc=qacc, M=M, qfrc_bias=c, qfrc_smooth=f, qfrc_constraint=J.T @ force if len(force) else np.zeros(dm.nv), ncon=len(cons),
                nefc=J.shape[0], ne=ne, efc_D=D, efc_aref=aref, actuator_force=frc, con_dist=np.array([k["dist"] for k in cons]), act_dot=act_dot,
                dfdv=dfdv)


def integrate_pos(dm, qpos, vel, h):
    q = np.array(qpos, float)
    for j in range(dm.njnt):
        jt, qa, da = int(dm.jnt_type[j]), int(dm.jnt_qposadr[j]), int(dm.jnt_dofadr[j])
        if jt in (SLIDE, HINGE):
            q[qa] += h * vel[da]
            continue
        if jt == FREE:
            q[qa:qa + 3] += h * vel[da:da + 3]
            qa, da = qa + 3, da + 3
        w = vel[da:da + 3]
        quat = q[qa:qa + 4] / np.linalg.norm(q[qa:qa + 4])
        ang = h * np.linalg.norm(w)
        if ang > 0:
            ax = w / np.linalg.norm(w)
            quat = quat_mul(quat, np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * ax]))
        q[qa:qa + 4] = quat
    return q


def step(dm: DenseModel, qpos, qvel, ctrl, qfrc_applied=None, xfrc_applied=None, act=None, mocap=None, eq_active=None, warmstart=None):
    """One mj_step. Returns dict(qpos, qvel, act, qacc, ncon, nefc, ...) - qacc is the forward's (pre-integration) acceleration."""
    h = dm.timestep
    qpos, qvel = np.asarray(qpos, float), np.asarray(qvel, float)
    act = np.zeros(dm.na) if act is None else np.asarray(act, float)
    f0 = forward(dm, qpos, qvel, ctrl, qfrc_applied, xfrc_applied, act, mocap, eq_active, warmstart)
    if dm.integrator in (0, 3):
        qacc = f0["qacc"]
        if dm.integrator == 3:      # implicitfast: (M - h d qfrc_smooth / d qvel) qacc' = M qacc, derivative = -damping + actuator term
            qacc = np.linalg.solve(f0["M"] + h * np.diag(dm.dof_damping - f0["dfdv"]), f0["qfrc_smooth"] + f0["qfrc_constraint"])
        elif np.any(dm.dof_damping > 0) and not dm.dis("eulerdamp"):
            qacc = np.linalg.solve(f0["M"] + h * np.diag(dm.dof_damping), f0["qfrc_smooth"] + f0["qfrc_constraint"])
        v1 = qvel + h * qacc
        out = dict(f0, qpos=integrate_pos(dm, qpos, v1, h), qvel=v1, act=next_activation(dm, act, f0["act_dot"]))
    else:                                                               # RK4, classic tableau
        A, Bw = [0.5, 0.5, 1.0], [1 / 6, 1 / 3, 1 / 3, 1 / 6]
        F = [(qvel, f0["qacc"], f0["act_dot"])]
        last = f0
        for i in range(3):
            qi = integrate_pos(dm, qpos, F[i][0], A[i] * h)
            vi = qvel + A[i] * h * F[i][1]
            ai = act + A[i] * h * F[i][2]
            last = forward(dm, qi, vi, ctrl, qfrc_applied, xfrc_applied, ai, mocap, eq_active)
            F.append((vi, last["qacc"], last["act_dot"]))
        sv = sum(w * Fi[0] for w, Fi in zip(Bw, F))
        sa = sum(w * Fi[1] for w, Fi in zip(Bw, F))
        sd = sum(w * Fi[2] for w, Fi in zip(Bw, F))
        out = dict(last, qpos=integrate_pos(dm, qpos, sv, h), qvel=qvel + h * sa, act=next_activation(dm, act, sd))   # derived fields: those of the 4th stage, as in mjData
    return out
