"""numpy/scipy twin of selected pieces of the path — an independent statement of the maths (no float32 tricks, library
linear algebra) used to cross-check the C++ oracle. Not the oracle and not the product."""
import numpy as np


def skew(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def distort(model, d, x, y):
    if model == 0:
        r2 = x * x + y * y
        r4 = r2 * r2
        x1 = x * (1 + d[4] * r2 + d[5] * r4) + 2 * d[6] * x * y + d[7] * (r2 + 2 * x * x)
        y1 = y * (1 + d[4] * r2 + d[5] * r4) + d[6] * (r2 + 2 * y * y) + 2 * d[7] * x * y
    else:
        r = np.sqrt(x * x + y * y)
        th = np.arctan(r)
        thd = th + d[4] * th**3 + d[5] * th**5 + d[6] * th**7 + d[7] * th**9
        c = thd / r if r > 1e-8 else 1.0
        x1, y1 = x * c, y * c
    return np.array([d[0] * x1 + d[2], d[1] * y1 + d[3]])


def project(frame, cam, cl, p_FinG, R=None, p=None, camR=None, camp=None, intr=None):
    """double-precision measurement function h(x): pixel of p_FinG seen by camera `cam` at clone `cl`."""
    R = frame.clone_R[cl].reshape(3, 3) if R is None else R
    p = frame.clone_p[cl] if p is None else p
    camR = frame.cam_R[cam].reshape(3, 3) if camR is None else camR
    camp = frame.cam_p[cam] if camp is None else camp
    intr = frame.cam_intr[cam] if intr is None else intr
    pc = camR @ (R @ (p_FinG - p)) + camp
    return distort(int(frame.cam_model[cam]), intr, pc[0] / pc[2], pc[1] / pc[2])


def exp_so3(w):
    th = np.linalg.norm(w)
    K = skew(w)
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th**2 * K @ K


def ekf_update(P, cols, H, res, Rdiag):
    """textbook EKF update on the sub-block: returns P+, dx."""
    Pc = P[:, cols]
    S = H @ P[np.ix_(cols, cols)] @ H.T + np.diag(Rdiag)
    K = np.linalg.solve(S, (Pc @ H.T).T).T
    Pn = P - K @ (Pc @ H.T).T
    Pn = np.triu(Pn) + np.triu(Pn, 1).T
    return Pn, K @ res


def triangulate_linear(frame, feats, f, anchor_cam, anchor_clone):
    """least-squares intersection of bearing rays in the anchor frame (same normal equations, numpy solve)."""
    def campose(cam, cl):
        R = frame.cam_R[cam].reshape(3, 3) @ frame.clone_R[cl].reshape(3, 3)
        return R, frame.clone_p[cl] - R.T @ frame.cam_p[cam]
    RA, pA = campose(anchor_cam, anchor_clone)
    A = np.zeros((3, 3))
    b = np.zeros(3)
    for i in range(feats.meas_off[f], feats.meas_off[f + 1]):
        Rc, pc = campose(int(feats.cam[i]), int(feats.clone[i]))
        Rac = Rc @ RA.T
        t = RA @ (pc - pA)
        bi = Rac.T @ np.array([feats.uvn[i, 0], feats.uvn[i, 1], 1.0], dtype=np.float64)
        bi /= np.linalg.norm(bi)
        Ai = np.eye(3) - np.outer(bi, bi)
        A += Ai
        b += Ai @ t
    pf = np.linalg.solve(A, b)
    return pf, RA.T @ pf + pA, np.linalg.cond(A)
