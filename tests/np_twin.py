"""Independent numpy twin of the dynamics / residual building blocks (test infrastructure).

Deliberately written differently from oracle/agx_oracle.cpp and from the CUDA kernels: body-frame
recursive Newton-Euler with 6x6 dense spatial matrices, and *complex-step differentiation* for every
derivative (exact to rounding for analytic functions), so that an error in the analytic
world-frame derivative algorithm cannot hide in both implementations.
"""
import numpy as np
import scipy.linalg


def skew(a):
    return np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=np.result_type(a, float))


def rot(axis, q):
    K = skew(np.asarray(axis, dtype=float))
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * (K @ K)


def joint_placement(t, i, q):
    """(R, p) of joint frame i in its parent body frame for coordinate q (complex-safe)."""
    if t.jtype[i] == 0:
        return t.placement_R[i] @ rot(t.axis[i], q), t.placement_p[i].astype(np.result_type(q, float))
    return t.placement_R[i].astype(np.result_type(q, float)), t.placement_p[i] + t.placement_R[i] @ (t.axis[i] * q)


def motion_xform_inv(R, p):
    """6x6 matrix taking a parent-frame motion [lin; ang] to the child frame."""
    X = np.zeros((6, 6), dtype=np.result_type(R, p))
    X[:3, :3] = R.T
    X[:3, 3:] = -R.T @ skew(p)
    X[3:, 3:] = R.T
    return X


def spatial_inertia(mass, c, I):
    Y = np.zeros((6, 6))
    C = skew(c)
    Y[:3, :3] = mass * np.eye(3)
    Y[:3, 3:] = -mass * C
    Y[3:, :3] = mass * C
    Y[3:, 3:] = I - mass * C @ C
    return Y


def crm(v):
    X = np.zeros((6, 6), dtype=v.dtype)
    X[:3, :3] = skew(v[3:])
    X[:3, 3:] = skew(v[:3])
    X[3:, 3:] = skew(v[3:])
    return X


def rnea(t, q, v, a):
    """tau = M(q) a + b(q, v), body-frame recursion; works on complex inputs."""
    nv = t.nv
    dt = np.result_type(q, v, a, float)
    S, Xi, vs, as_, fs = [], [], [], [], []
    for i in range(nv):
        R, p = joint_placement(t, i, q[i])
        X = motion_xform_inv(R, p)
        s = np.zeros(6)
        if t.jtype[i] == 0:
            s[3:] = t.axis[i]
        else:
            s[:3] = t.axis[i]
        par = t.parent[i]
        vp = vs[par] if par >= 0 else np.zeros(6, dtype=dt)
        ap = as_[par] if par >= 0 else np.concatenate([-t.gravity, np.zeros(3)]).astype(dt)
        vi = X @ vp + s * v[i]
        ai = X @ ap + s * a[i] + crm(vi) @ (s * v[i])
        Y = spatial_inertia(t.mass[i], t.com[i], t.inertia[i])
        fi = Y @ ai - crm(vi).T @ (Y @ vi)
        S.append(s); Xi.append(X); vs.append(vi); as_.append(ai); fs.append(fi)
    tau = np.zeros(nv, dtype=dt)
    for i in reversed(range(nv)):
        tau[i] = S[i] @ fs[i]
        if t.parent[i] >= 0:
            fs[t.parent[i]] = fs[t.parent[i]] + Xi[i].T @ fs[i]
    return tau


def mass_matrix(t, q):
    nv = t.nv
    z = np.zeros(nv)
    b0 = rnea_nograv(t, q, z, z)
    return np.stack([rnea_nograv(t, q, z, e) - b0 for e in np.eye(nv)], axis=1)


def rnea_nograv(t, q, v, a):
    import dataclasses

    return rnea(dataclasses.replace(t, gravity=np.zeros(3)), q, v, a)


def forward_dynamics(t, q, v, u):
    M = mass_matrix(t, q) + np.diag(t.armature)
    b = rnea(t, q, v, np.zeros(t.nv))
    return np.linalg.solve(M, u - b)


def complex_step_jac(f, x, h=1e-30):
    x = np.asarray(x, dtype=float)
    cols = []
    for i in range(x.size):
        xc = x.astype(complex)
        xc[i] += 1j * h
        cols.append(np.imag(f(xc)) / h)
    return np.stack(cols, axis=-1)


def frame_placement(t, q):
    """World placement (4x4 homogeneous) of the task frame."""
    dt = np.result_type(q, float)
    Ms = []
    for i in range(t.nv):
        R, p = joint_placement(t, i, q[i])
        M = np.eye(4, dtype=dt)
        M[:3, :3], M[:3, 3] = R, p
        Ms.append(M if t.parent[i] < 0 else Ms[t.parent[i]] @ M)
    par, fR, fp = t.frames[t.frame_name]
    F = np.eye(4)
    F[:3, :3], F[:3, 3] = fR, fp
    return Ms[par] @ F


def log6_logm(M):
    """se(3) logarithm [lin; ang] through scipy's matrix logarithm (independent of Pinocchio's formulas)."""
    L = np.real(scipy.linalg.logm(M))
    return np.array([L[0, 3], L[1, 3], L[2, 3], L[2, 1], L[0, 2], L[1, 0]])


def exp6(xi):
    X = np.zeros((4, 4))
    X[:3, :3] = skew(xi[3:])
    X[:3, 3] = xi[:3]
    return scipy.linalg.expm(X)
