"""Numerical-analysis yardsticks for the fp32 throughput mode (used by tests/test_gpu_fp32.py and tools/fp32_study.py).

The fp32 step cannot meet a 1e-4 FORWARD error on qacc for contact-rich, ill-conditioned models: any fp32 evaluation of
qacc = argmin 1/2 (a-a0)'M(a-a0) + s(Ja - aref) carries cond(H) * eps error. What a correct fp32 kernel CAN guarantee is a small
BACKWARD error: its qacc satisfies the fp64 equations up to a residual of a few fp32 ulps relative to the terms that make it
up. backward_error() measures exactly that against the fp64 oracle's M, J, D, aref, qfrc_smooth at the same state."""
import numpy as np


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


def backward_error(model, od, a):
    """Normwise backward error of `a` as a solution of the stationarity condition  M a - qfrc_smooth - J' f(a) = 0  with
    f_r = -D_r min(0, J_r a - aref_r)  (SURVEY A.11), all quantities from the fp64 oracle `od` after forward():
        eta = |r|_inf / (|M|_inf |a|_inf + |qfrc_smooth|_inf + |J'|_inf |f|_inf)
    and cond_inf(H) of the active-set Hessian H = M + J' D_active J that maps a residual into an error in qacc."""
    nv, nefc = model.nv, od.int("nefc")
    M = dense_M(model, od.field("qM"))
    fs = od.field("qfrc_smooth")
    a = np.asarray(a, dtype=np.float64)
    if nefc:
        J = od.field("efc_J")[:nefc * nv].reshape(nefc, nv)
        D, aref = od.field("efc_D")[:nefc], od.field("efc_aref")[:nefc]
        jar = J @ a - aref
        f = np.where(jar < 0, -D * jar, 0.0)
        r = M @ a - fs - J.T @ f
        act = jar < 0
        H = M + (J[act].T * D[act]) @ J[act]
        scale = np.abs(M).sum(1).max() * np.abs(a).max() + np.abs(fs).max() + np.abs(J.T).sum(1).max() * np.abs(f).max()
    else:
        r = M @ a - fs
        H = M
        scale = np.abs(M).sum(1).max() * np.abs(a).max() + np.abs(fs).max()
    return float(np.abs(r).max() / max(scale, 1e-300)), float(np.linalg.cond(H, np.inf))
