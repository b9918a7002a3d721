"""ORACLE (test infrastructure, not product code) — Lie-group maps restated on the CPU in NumPy.

PARITY UNPINNED: the arithmetic of iSLAM's back-end lives in PyPose, which is not vendored in
/root/reference, not pinned in its environment.yml and not installed here.  This file restates the
published PyPose 0.6.x semantics used at the reference call sites
  pvgo.py:36-39,45-48,72,116-118   imu_integrator.py:146,151   Datasets/transformation.py:72-124
and is pinned only by the known-answer tests in tests/test_oracle_lie.py (SciPy Rotation / expm).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Conventions:  SE3 = [t(3), q=(x,y,z,w)],  se3 = [tau(3), phi(3)],  SO3 = (x,y,z,w).
All functions are batched over leading dims and keep the dtype of their input (float32 or float64).
"""
import numpy as np

# below these angles the closed forms lose digits to cancellation; use 4-term Taylor series instead
_EPS = {np.dtype(np.float32): 0.5, np.dtype(np.float64): 0.05}


def _eps(x):
    return _EPS[np.dtype(x.dtype)]


def skew(v):
    z = np.zeros_like(v[..., 0])
    return np.stack([
        np.stack([z, -v[..., 2], v[..., 1]], -1),
        np.stack([v[..., 2], z, -v[..., 0]], -1),
        np.stack([-v[..., 1], v[..., 0], z], -1)], -2)


# ----------------------------------------------------------------------------- SO3
def so3_exp(phi):
    """so3.Exp: q = [sin(th/2)/th * phi, cos(th/2)]  (A.2)."""
    th2 = np.sum(phi * phi, -1, keepdims=True)
    th = np.sqrt(th2)
    small = th < _eps(phi)
    ths = np.where(small, 1, th)
    k = np.where(small, 0.5 - th2 / 48.0 + th2 * th2 / 3840.0 - th2 ** 3 / 645120.0, np.sin(0.5 * ths) / ths)
    w = np.cos(0.5 * th)
    return np.concatenate([k * phi, w], -1).astype(phi.dtype)


def so3_log(q):
    """SO3.Log: phi = 2*atan(|v|/w)/|v| * v ; q and -q give the same phi (A.2)."""
    v, w = q[..., :3], q[..., 3:4]
    n2 = np.sum(v * v, -1, keepdims=True)
    n = np.sqrt(n2)
    small = n < 1e-6
    ns = np.where(small, 1, n)
    ws = np.where(np.abs(w) < 1e-30, 1e-30, w)
    f = np.where(small, 2.0 / ws - (2.0 / 3.0) * n2 / (ws * ws * ws), 2.0 * np.arctan(ns / ws) / ns)
    # w == 0 exactly: rotation by pi
    f = np.where(np.abs(w) < 1e-30, np.pi / ns, f)
    return (f * v).astype(q.dtype)


def so3_inv(q):
    return q * np.array([-1, -1, -1, 1], q.dtype)


def so3_mul(a, b):
    ax, ay, az, aw = (a[..., i] for i in range(4))
    bx, by, bz, bw = (b[..., i] for i in range(4))
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by - ax * bz + ay * bw + az * bx,
        aw * bz + ax * by - ay * bx + az * bw,
        aw * bw - ax * bx - ay * by - az * bz], -1)


def so3_act(q, p):
    """R(q) p  via  p + 2 w (v x p) + 2 v x (v x p)."""
    v, w = q[..., :3], q[..., 3:4]
    t = 2.0 * np.cross(v, p)
    return p + w * t + np.cross(v, t)


def so3_matrix(q):
    x, y, z, w = (q[..., i] for i in range(4))
    return np.stack([
        np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)], -1),
        np.stack([2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)], -1),
        np.stack([2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1)], -2)


def so3_Jl(phi):
    """Jl = I + (1-cos)/th^2 K + (th-sin)/th^3 K^2  (A.2)."""
    K = skew(phi)
    th2 = np.sum(phi * phi, -1)[..., None, None]
    th = np.sqrt(th2)
    small = th < _eps(phi)
    ths = np.where(small, 1, th)
    c1 = np.where(small, 0.5 - th2 / 24.0 + th2 ** 2 / 720.0 - th2 ** 3 / 40320.0, (1 - np.cos(ths)) / (ths * ths))
    c2 = np.where(small, 1.0 / 6.0 - th2 / 120.0 + th2 ** 2 / 5040.0 - th2 ** 3 / 362880.0,
                  (ths - np.sin(ths)) / (ths ** 3))
    I = np.eye(3, dtype=phi.dtype)
    return (I + c1 * K + c2 * (K @ K)).astype(phi.dtype)


def so3_Jl_inv(phi):
    """Jl^-1 = I - K/2 + (1/th^2 - (1+cos)/(2 th sin)) K^2  (A.2; -> 1/12 as th -> 0)."""
    K = skew(phi)
    th2 = np.sum(phi * phi, -1)[..., None, None]
    th = np.sqrt(th2)
    small = th < _eps(phi)
    ths = np.where(small, 1, th)
    c = np.where(small, 1.0 / 12.0 + th2 / 720.0 + th2 ** 2 / 30240.0 + th2 ** 3 / 1209600.0,
                 1.0 / (ths * ths) - (1 + np.cos(ths)) / (2 * ths * np.sin(ths)))
    I = np.eye(3, dtype=phi.dtype)
    return (I - 0.5 * K + c * (K @ K)).astype(phi.dtype)


# ----------------------------------------------------------------------------- SE3
def se3_exp(xi):
    """se3.Exp: t = Jl(phi) tau, q = so3.Exp(phi)  (A.2)."""
    tau, phi = xi[..., :3], xi[..., 3:]
    t = (so3_Jl(phi) @ tau[..., None])[..., 0]
    return np.concatenate([t, so3_exp(phi)], -1).astype(xi.dtype)


def se3_log(X):
    """SE3.Log: phi = Log(q), tau = Jl^-1(phi) t  (A.2)."""
    t, q = X[..., :3], X[..., 3:]
    phi = so3_log(q)
    tau = (so3_Jl_inv(phi) @ t[..., None])[..., 0]
    return np.concatenate([tau, phi], -1).astype(X.dtype)


def se3_inv(X):
    t, q = X[..., :3], X[..., 3:]
    qi = so3_inv(q)
    return np.concatenate([-so3_act(qi, t), qi], -1)


def se3_mul(A, B):
    ta, qa = A[..., :3], A[..., 3:]
    tb, qb = B[..., :3], B[..., 3:]
    return np.concatenate([ta + so3_act(qa, tb), so3_mul(qa, qb)], -1)


def se3_act(X, p):
    return so3_act(X[..., 3:], p) + X[..., :3]


def se3_adj(X):
    """Ad(T) = [[R, [t]x R], [0, R]] in [tau, phi] ordering (A.2)."""
    R = so3_matrix(X[..., 3:])
    tR = skew(X[..., :3]) @ R
    Z = np.zeros_like(R)
    return np.concatenate([np.concatenate([R, tR], -1), np.concatenate([Z, R], -1)], -2)


def se3_Jl_inv(xi):
    """6x6 Jl^-1(xi) = [[Jl^-1, -Jl^-1 Q Jl^-1], [0, Jl^-1]] with Barfoot's Q (A.2)."""
    tau, phi = xi[..., :3], xi[..., 3:]
    T, K = skew(tau), skew(phi)
    th2 = np.sum(phi * phi, -1)[..., None, None]
    th = np.sqrt(th2)
    small = th < _eps(xi)
    ths = np.where(small, 1, th)
    s, c = np.sin(ths), np.cos(ths)
    c1 = np.where(small, 1.0 / 6.0 - th2 / 120.0 + th2 ** 2 / 5040.0 - th2 ** 3 / 362880.0, (ths - s) / ths ** 3)
    c2 = np.where(small, 1.0 / 24.0 - th2 / 720.0 + th2 ** 2 / 40320.0 - th2 ** 3 / 3628800.0,
                  (ths * ths + 2 * c - 2) / (2 * ths ** 4))
    c3 = np.where(small, 1.0 / 120.0 - th2 / 2520.0 + th2 ** 2 / 120960.0 - th2 ** 3 / 9979200.0,
                  (2 * ths - 3 * s + ths * c) / (2 * ths ** 5))
    KT, TK = K @ T, T @ K
    KTK = KT @ K
    Q = 0.5 * T + c1 * (KT + TK + KTK) + c2 * (K @ KT + TK @ K - 3 * KTK) + c3 * (KTK @ K + K @ KTK)
    Ji = so3_Jl_inv(phi)
    Z = np.zeros_like(Ji)
    B = -Ji @ Q @ Ji
    return np.concatenate([np.concatenate([Ji, B], -1), np.concatenate([Z, Ji], -1)], -2).astype(xi.dtype)


def se3_retract(X, delta):
    """LieTensor.add_: X <- Exp(delta[..., :6]) * X  (left perturbation, A.1)."""
    return se3_mul(se3_exp(delta[..., :6].astype(X.dtype)), X)


def quat_canon(q):
    """Flip sign so that w >= 0 (for comparisons only)."""
    return np.where(q[..., 3:4] < 0, -q, q)
