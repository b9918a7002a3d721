"""ORACLE (test infrastructure, not product code) — CPU restatement of iSLAM's PVGO back-end.

PARITY UNPINNED: PyPose (the library that holds the arithmetic) is absent from /root/reference and from
this image and the reference ships no tests; this restates SURVEY.md Appendix A.3-A.4 (PyPose 0.6.x
`pp.optim.LM.step`, `solver.Cholesky`, `strategy.TrustRegion`, `scheduler.StopOnPlateau`) for the problem
defined in /root/reference/pvgo.py.  Each function cites the reference lines it follows.

Two solvers with identical mathematics:
  * `DenseLM`   — literal: dense J (R x 10N, dead 7th pose column kept), dense block-diag W (R x R),
                  A = J^T W J, clamp, cumulative damping, torch cholesky_ex / cholesky_solve.
                  Only runnable for small N (C1; N <~ 500).
  * `SparseLM`  — same normal equations assembled block-sparse on 9N unknowns (the dead columns are
                  decoupled: diag clamped to 1e-4, gradient 0 => delta 0) and solved with a banded /
                  sparse Cholesky.  Runs C2-C4; this is the "CPU reference" timed beside GPU numbers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
import numpy as np
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import lie
from . import reproj_oracle


# --------------------------------------------------------------------------------------------------
# residuals  (pvgo.py:26-64)
# --------------------------------------------------------------------------------------------------
def residuals(nodes, vels, edges, poses, drots, dtrans, dvels, dts):
    """Returns (pgerr (E,6), adjvelerr (M,3), imuroterr (M,3), transvelerr (M,3)) — pvgo.py:36-51,64."""
    n1, n2 = nodes[edges[:, 0]], nodes[edges[:, 1]]
    err = lie.se3_mul(lie.se3_mul(lie.se3_inv(poses), lie.se3_inv(n1)), n2)          # pvgo.py:38
    pgerr = lie.se3_log(err)                                                          # pvgo.py:39
    adjvelerr = dvels - np.diff(vels, axis=0)                                         # pvgo.py:42
    r1, r2 = nodes[:-1, 3:], nodes[1:, 3:]
    rerr = lie.so3_mul(lie.so3_mul(lie.so3_inv(drots), lie.so3_inv(r1)), r2)          # pvgo.py:47
    imuroterr = lie.so3_log(rerr)                                                     # pvgo.py:48
    transvelerr = np.diff(nodes[:, :3], axis=0) - (vels[:-1] * dts[:, None] + dtrans)  # pvgo.py:51
    return pgerr, adjvelerr, imuroterr, transvelerr


def loss_of(res):
    """RobustModel.loss with the Trivial kernel: UNWEIGHTED sum of squares (A.4)."""
    return sum(float(np.sum(r.astype(np.float64) ** 2)) if r.dtype == np.float64
               else float(np.sum(r * r, dtype=r.dtype)) for r in res)


def jacobian_blocks(nodes, vels, edges, poses, drots, pgerr, imuroterr):
    """Closed-form blocks PyPose's autograd yields (A.3).

    Jvo (E,6,6)  = Jl^-1(r) Ad(Z^-1 Xi^-1)      d r_e / d delta_j ;  d/d delta_i = -Jvo
    Jrot (M,3,3) = Jl^-1(r) dR^T Ri^T            d r_i / d phi_{i+1};  d/d phi_i  = -Jrot
    """
    n1 = nodes[edges[:, 0]]
    A = lie.se3_mul(lie.se3_inv(poses), lie.se3_inv(n1))
    Jvo = lie.se3_Jl_inv(pgerr) @ lie.se3_adj(A)
    Rq = lie.so3_mul(lie.so3_inv(drots), lie.so3_inv(nodes[:-1, 3:]))
    Jrot = lie.so3_Jl_inv(imuroterr) @ lie.so3_matrix(Rq)
    return Jvo.astype(nodes.dtype), Jrot.astype(nodes.dtype)


def info_scalars(loss_weight):
    """pvgo.py:125-129: VO info = w0^2 (trans AND rot), dv = w1^2, drot = w2^2, transvel = w3^2."""
    w = [float(x) for x in loss_weight]
    return w[0] ** 2, w[1] ** 2, w[2] ** 2, w[3] ** 2


# --------------------------------------------------------------------------------------------------
# shared LM controller (A.4): trust region + accept/reject, independent of how A x = b is solved
# --------------------------------------------------------------------------------------------------
class _LMBase:
    def __init__(self, g, dtype=np.float64, radius=None, lm_min=1e-4, lm_max=1e32, reject=16,
                 rollback='minus_d'):
        self.dtype = np.dtype(dtype)
        c = lambda a: np.ascontiguousarray(a, dtype=self.dtype)
        self.nodes, self.vels = c(g.init_nodes).copy(), c(g.init_vels).copy()          # pvgo.py:20-21
        self.edges = np.asarray(g.links, dtype=np.int64)
        self.poses, self.drots = c(g.vo_motions), c(g.imu_drots)
        self.dtrans, self.dvels, self.dts = c(g.imu_dtrans), c(g.imu_dvels), c(g.dts)
        self.N, self.E, self.M = self.nodes.shape[0], self.edges.shape[0], self.nodes.shape[0] - 1
        assert self.dts.shape[0] == self.M                                              # pvgo.py:155 / 8a a5
        self.w = info_scalars(g.loss_weight)
        radius = g.radius if radius is None else radius
        # TrustRegion(radius): high=.5 low=1e-3 up=2 down=.5 factor=.5 min=1e-6 max=1e16   (A.4)
        self.tr = dict(radius=float(radius), high=0.5, low=1e-3, up=2.0, down=0.5, factor=0.5,
                       min=1e-6, max=1e16, down0=0.5)
        self.damping = 1.0 / float(radius)
        self.lm_min, self.lm_max, self.reject = lm_min, lm_max, reject
        self.rollback = rollback
        self.loss = None
        self.last = None
        self.reject_count = 0
        self.history = []
        # optional 5th residual group (pvgo.py:53-61): g.extra['reproj'] = dict(point3d, target, K, rgb2imu, N, weight)
        self.rp = (getattr(g, 'extra', None) or {}).get('reproj')
        self.w4 = (float(self.rp['weight']) / int(self.rp['N'])) ** 2 if self.rp is not None else 0.0     # pvgo.py:131

    # -- model --------------------------------------------------------------------------------------
    def _res(self):
        res = residuals(self.nodes, self.vels, self.edges, self.poses, self.drots, self.dtrans,
                        self.dvels, self.dts)
        if self.rp is not None:
            res = res + (reproj_oracle.residual(self.nodes, self.rp),)       # pvgo.py:58-61
        return res

    def _update(self, dn, dv, sign=1.0):
        """update_parameter: nodes <- Exp(d) nodes ; vels <- vels + d  (A.1/A.4)."""
        self.nodes = lie.se3_retract(self.nodes, (sign * dn).astype(self.dtype))
        self.vels = (self.vels + sign * dv).astype(self.dtype)

    def _tr_update(self, last, loss, denom):
        """TrustRegion.update (A.4).  quality = (last - loss) / -((JD)^T (2R + JD)), unweighted."""
        tr = self.tr
        quality = (last - loss) / denom if denom != 0 else np.inf * np.sign(last - loss)
        tr['radius'] = 1.0 / self.damping
        if quality > tr['high']:
            tr['radius'] *= tr['up']
            tr['down'] = tr['down0']
        elif quality > tr['low']:
            tr['down'] = tr['down0']
        else:
            tr['radius'] *= tr['down']
            tr['down'] *= tr['factor']
        tr['down'] = max(tr['min'], min(tr['down'], tr['max']))
        tr['radius'] = max(tr['min'], min(tr['radius'], tr['max']))
        self.damping = 1.0 / tr['radius']
        return quality

    # -- one optimizer.step (A.4) ---------------------------------------------------------------------
    def step(self):
        res = self._res()
        lin = self._linearize(res)
        if self.loss is None:
            self.loss = loss_of(res)
        self.last = self.loss
        self.reject_count = 0
        tries = 0
        while self.last <= self.loss:
            lin = self._damp(lin, self.damping)
            ok, dn, dv = self._solve(lin)
            if not ok:
                print('Linear solver failed. Breaking optimization step...')
                break
            backup = (self.nodes.copy(), self.vels.copy())
            self._update(dn, dv)
            self.loss = loss_of(self._res())
            denom = self._quality_denominator(lin, res, dn, dv)
            q = self._tr_update(self.last, self.loss, denom)
            tries += 1
            if self.last < self.loss and self.reject_count < self.reject:
                if self.rollback == 'minus_d':
                    self._update(dn, dv, sign=-1.0)            # PyPose applies -D (A.4 last note)
                else:
                    self.nodes, self.vels = backup
                self.loss = self.last
                self.reject_count += 1
            else:
                break
        self.history.append(dict(loss=self.loss, last=self.last, rejects=self.reject_count,
                                 damping=self.damping))
        return self.loss

    def run(self, steps=None, scheduler=True, max_steps=10, patience=3, decreasing=1e-3):
        """pvgo.py:172-180.  steps=k => k fixed optimizer.step calls (scheduler bypassed)."""
        if steps is not None:
            for _ in range(steps):
                self.step()
            return self
        n, pc = 0, 0
        while True:                                                     # StopOnPlateau (A.4)
            loss = self.step()
            n += 1
            stop = n >= max_steps
            pc = pc + 1 if (self.last - loss) < decreasing else 0
            stop = stop or pc >= patience or self.reject_count >= self.reject
            if stop:
                break
        return self

    # -- outputs (pvgo.py:67-78, 114-119, 186-197) -----------------------------------------------------
    def vo_loss(self, vo_motions=None):
        P = self.poses if vo_motions is None else np.asarray(vo_motions, self.dtype)
        n1, n2 = self.nodes[self.edges[:, 0]], self.nodes[self.edges[:, 1]]
        e = lie.se3_log(lie.se3_mul(lie.se3_mul(lie.se3_inv(P), lie.se3_inv(n1)), n2))
        return np.sum(e[:, :3] ** 2, 1), np.sum(e[:, 3:] ** 2, 1)

    def imu_loss(self):
        _, adj, rot, _ = self._res()
        return np.sum(adj ** 2, 1), np.sum(rot ** 2, 1)

    def aligned(self, target):
        return align_to(self.nodes, self.vels, np.asarray(target, self.dtype))


def align_to(nodes, vels, target, idx=0):
    """pvgo.py:114-119: nodes <- T X0^-1 nodes ; vels <- R_T R0^-1 vels."""
    src = nodes[idx]
    T = lie.se3_mul(target, lie.se3_inv(src))
    out_n = lie.se3_mul(T[None], nodes)
    q = lie.so3_mul(target[3:], lie.so3_inv(src[3:]))
    out_v = lie.so3_act(q[None], vels)
    return out_n, out_v


def vo_loss_grad(nodes, edges, P):
    """d(trans_loss_e)/dP_e and d(rot_loss_e)/dP_e as left-tangent 6-vectors (A.3 last row):
    e = Log(P^-1 n1^-1 n2);  de/d(delta_P) = -Jl^-1(e) Ad(P^-1).  Returns (gt (E,6), gr (E,6))."""
    n1, n2 = nodes[edges[:, 0]], nodes[edges[:, 1]]
    Pi = lie.se3_inv(P)
    e = lie.se3_log(lie.se3_mul(lie.se3_mul(Pi, lie.se3_inv(n1)), n2))
    J = -lie.se3_Jl_inv(e) @ lie.se3_adj(Pi)
    et = e.copy(); et[:, 3:] = 0
    er = e.copy(); er[:, :3] = 0
    gt = 2 * np.einsum('ek,ekj->ej', et, J)
    gr = 2 * np.einsum('ek,ekj->ej', er, J)
    return gt, gr


# --------------------------------------------------------------------------------------------------
# literal dense LM
# --------------------------------------------------------------------------------------------------
class DenseLM(_LMBase):
    """PyPose's dense algorithm, column layout [nodes 7N | vels 3N], rows [6E | 3M | 3M | 3M] (A.3/A.4)."""

    def dense_J(self, res):
        if self.rp is not None:
            raise NotImplementedError('the dense literal oracle covers the four residual groups train.py uses; '
                                      'the reprojection factor is in SparseLM')
        N, E, M = self.N, self.E, self.M
        Jvo, Jrot = jacobian_blocks(self.nodes, self.vels, self.edges, self.poses, self.drots, res[0], res[2])
        R = 6 * E + 9 * M
        J = np.zeros((R, 10 * N), self.dtype)
        for e in range(E):
            i, j = self.edges[e]
            J[6 * e:6 * e + 6, 7 * j:7 * j + 6] += Jvo[e]
            J[6 * e:6 * e + 6, 7 * i:7 * i + 6] -= Jvo[e]
        o = 6 * E
        I3 = np.eye(3, dtype=self.dtype)
        for i in range(M):                                   # adjvelerr = dv - (v_{i+1} - v_i)
            r = o + 3 * i
            J[r:r + 3, 7 * N + 3 * i:7 * N + 3 * i + 3] += I3
            J[r:r + 3, 7 * N + 3 * (i + 1):7 * N + 3 * (i + 1) + 3] -= I3
        o += 3 * M
        for i in range(M):                                   # imuroterr
            r = o + 3 * i
            J[r:r + 3, 7 * (i + 1) + 3:7 * (i + 1) + 6] += Jrot[i]
            J[r:r + 3, 7 * i + 3:7 * i + 6] -= Jrot[i]
        o += 3 * M
        for i in range(M):                                   # transvelerr (phi columns ZERO — A.3 quirk)
            r = o + 3 * i
            J[r:r + 3, 7 * (i + 1):7 * (i + 1) + 3] += I3
            J[r:r + 3, 7 * i:7 * i + 3] -= I3
            J[r:r + 3, 7 * N + 3 * i:7 * N + 3 * i + 3] -= self.dts[i] * I3
        return J

    def _linearize(self, res):
        import torch
        J = self.dense_J(res)
        Rv = np.concatenate([r.reshape(-1) for r in res]).astype(self.dtype)
        w0, w1, w2, w3 = self.w
        wdiag = np.concatenate([np.full(6 * self.E, w0), np.full(3 * self.M, w1), np.full(3 * self.M, w2),
                                np.full(3 * self.M, w3)]).astype(self.dtype)
        Jt, Rt = torch.from_numpy(J), torch.from_numpy(Rv)
        W = torch.diag(torch.from_numpy(wdiag))             # torch.block_diag of diagonal blocks == diag
        J_T = Jt.T @ W
        A = J_T @ Jt
        A.diagonal().clamp_(self.lm_min, self.lm_max)
        b = -(J_T @ Rt.view(-1, 1))
        return dict(J=Jt, R=Rt, A=A, b=b)

    def _damp(self, lin, damping):
        lin['A'].diagonal().add_(lin['A'].diagonal() * damping)      # cumulative (A.4)
        return lin

    def _solve(self, lin):
        import torch
        L, info = torch.linalg.cholesky_ex(lin['A'])
        if int(info) != 0 or bool(torch.isnan(L).any()):
            return False, None, None
        D = torch.cholesky_solve(lin['b'], L).view(-1).numpy()
        N = self.N
        lin['D'] = D
        return True, D[:7 * N].reshape(N, 7)[:, :6].copy(), D[7 * N:].reshape(N, 3).copy()

    def _quality_denominator(self, lin, res, dn, dv):
        import torch
        JD = lin['J'] @ torch.from_numpy(lin['D']).view(-1, 1)
        return float(-(JD.T @ (2 * lin['R'].view(-1, 1) + JD)))


# --------------------------------------------------------------------------------------------------
# sparse twin: same normal equations on 9N unknowns, node-major layout [tau, phi, v] per node
# --------------------------------------------------------------------------------------------------
class SparseLM(_LMBase):
    def __init__(self, g, dtype=np.float64, solver='auto', **kw):
        super().__init__(g, dtype=dtype, **kw)
        self.solver = solver
        bw = int(np.max(np.abs(self.edges[:, 1] - self.edges[:, 0]))) if self.E else 1
        self.band_nodes = max(bw, 1)

    def assemble(self, res):
        """Block-sparse H = J^T W J (before clamp/damp) and g = J^T W r as COO, plus J pieces."""
        N, E, M = self.N, self.E, self.M
        w0, w1, w2, w3 = self.w
        Jvo, Jrot = jacobian_blocks(self.nodes, self.vels, self.edges, self.poses, self.drots, res[0], res[2])
        dt = self.dts
        S = w0 * np.einsum('eki,ekj->eij', Jvo, Jvo)          # (E,6,6)
        q = w0 * np.einsum('eki,ek->ei', Jvo, res[0])         # (E,6)   J^T W r at node j; -q at node i
        Sr = w2 * np.einsum('mki,mkj->mij', Jrot, Jrot)       # (M,3,3)
        qr = w2 * np.einsum('mki,mk->mi', Jrot, res[2])
        rows, cols, vals = [], [], []

        def add_block(bi, bj, ro, co, blk):
            """blk: (K,a,b) added at node-block (bi,bj) with in-block offsets (ro,co)."""
            K, a, b = blk.shape
            r = (9 * bi + ro)[:, None, None] + np.arange(a)[None, :, None]
            c = (9 * bj + co)[:, None, None] + np.arange(b)[None, None, :]
            rows.append(np.broadcast_to(r, blk.shape).ravel())
            cols.append(np.broadcast_to(c, blk.shape).ravel())
            vals.append(blk.ravel())

        i, j = self.edges[:, 0], self.edges[:, 1]
        add_block(i, i, 0, 0, S); add_block(j, j, 0, 0, S)
        add_block(i, j, 0, 0, -S); add_block(j, i, 0, 0, -S)
        a = np.arange(M); b = a + 1
        I3 = np.broadcast_to(np.eye(3, dtype=self.dtype), (M, 3, 3))
        # delta-velocity: J(v_i)=+I, J(v_{i+1})=-I
        add_block(a, a, 6, 6, w1 * I3); add_block(b, b, 6, 6, w1 * I3)
        add_block(a, b, 6, 6, -w1 * I3); add_block(b, a, 6, 6, -w1 * I3)
        # imu rotation: J(phi_{i+1})=Jrot, J(phi_i)=-Jrot
        add_block(a, a, 3, 3, Sr); add_block(b, b, 3, 3, Sr)
        add_block(a, b, 3, 3, -Sr); add_block(b, a, 3, 3, -Sr)
        # trans-vel: J(tau_{i+1})=+I, J(tau_i)=-I, J(v_i)=-dt I
        dI = dt[:, None, None] * I3
        add_block(a, a, 0, 0, w3 * I3); add_block(b, b, 0, 0, w3 * I3)
        add_block(a, b, 0, 0, -w3 * I3); add_block(b, a, 0, 0, -w3 * I3)
        add_block(a, a, 0, 6, w3 * dI); add_block(a, a, 6, 0, w3 * dI)
        add_block(a, a, 6, 6, w3 * dI * dt[:, None, None])
        add_block(b, a, 0, 6, -w3 * dI); add_block(a, b, 6, 0, -w3 * dI)
        Jrp = None
        if self.rp is not None:
            # reprojection: J(delta_i) = +Jrp, J(delta_{i+1}) = -Jrp (pair 0: zero), information w4 I
            Jrp = reproj_oracle.jacobian(self.nodes, self.rp).astype(self.dtype)          # (M, 2Np, 6)
            Sp = self.w4 * np.einsum('mki,mkj->mij', Jrp, Jrp)
            add_block(a, a, 0, 0, Sp); add_block(b, b, 0, 0, Sp)
            add_block(a, b, 0, 0, -Sp); add_block(b, a, 0, 0, -Sp)
        H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(9 * N, 9 * N), dtype=self.dtype).tocsc()
        g = np.zeros((N, 9), self.dtype)
        if Jrp is not None:
            qp = self.w4 * np.einsum('mki,mk->mi', Jrp, res[4])
            g[:-1, 0:6] += qp; g[1:, 0:6] -= qp
        self._Jrp = Jrp
        np.add.at(g[:, 0:6], j, q); np.add.at(g[:, 0:6], i, -q)
        g[1:, 3:6] += qr; g[:-1, 3:6] -= qr
        g[:-1, 6:9] += w1 * res[1]; g[1:, 6:9] -= w1 * res[1]
        g[1:, 0:3] += w3 * res[3]; g[:-1, 0:3] -= w3 * res[3]
        g[:-1, 6:9] -= w3 * dt[:, None] * res[3]
        return H, g, Jvo, Jrot

    def _linearize(self, res):
        H, g, Jvo, Jrot = self.assemble(res)
        d = H.diagonal().copy()
        d = np.clip(d, self.lm_min, self.lm_max).astype(self.dtype)
        return dict(H=H, diag0=H.diagonal().copy(), diag=d, g=g, Jvo=Jvo, Jrot=Jrot, Jrp=self._Jrp)

    def _damp(self, lin, damping):
        lin['diag'] = (lin['diag'] + lin['diag'] * self.dtype.type(damping)).astype(self.dtype)
        return lin

    def _solve(self, lin):
        n = 9 * self.N
        A = (lin['H'] + sp.diags(lin['diag'] - lin['diag0'], format='csc')).tocsc()
        b = -lin['g'].reshape(-1)
        use_band = self.solver == 'band' or (self.solver == 'auto' and self.band_nodes <= 16)
        try:
            if use_band:
                u = 9 * self.band_nodes + 8
                ab = np.zeros((u + 1, n), self.dtype)
                C = A.tocoo()
                m = C.row >= C.col
                ab[(C.row - C.col)[m], C.col[m]] = C.data[m]
                cb = scipy.linalg.cholesky_banded(ab, lower=True, check_finite=False)
                D = scipy.linalg.cho_solve_banded((cb, True), b, check_finite=False)
            else:
                lu = spla.splu(A, permc_spec='MMD_AT_PLUS_A', diag_pivot_thresh=0.0,
                               options=dict(SymmetricMode=True))
                D = lu.solve(b)
        except Exception as ex:           # non-PD pivot => PyPose prints and breaks the step
            print(ex)
            return False, None, None
        if not np.all(np.isfinite(D)):
            return False, None, None
        D = D.astype(self.dtype).reshape(self.N, 9)
        lin['D'] = D
        return True, D[:, :6].copy(), D[:, 6:].copy()

    def _quality_denominator(self, lin, res, dn, dv):
        """-(JD)^T(2R+JD) evaluated factor by factor with the linearisation-point J and r (unweighted)."""
        i, j = self.edges[:, 0], self.edges[:, 1]
        jd0 = np.einsum('ekj,ej->ek', lin['Jvo'], dn[j] - dn[i])
        jd1 = dv[:-1] - dv[1:]
        jd2 = np.einsum('mkj,mj->mk', lin['Jrot'], dn[1:, 3:6] - dn[:-1, 3:6])
        jd3 = dn[1:, 0:3] - dn[:-1, 0:3] - self.dts[:, None] * dv[:-1]
        tot = 0.0
        for jd, r in zip((jd0, jd1, jd2, jd3), res):
            tot += float(np.sum(jd * (2 * r + jd)))
        if lin.get('Jrp') is not None:
            jd4 = np.einsum('mkj,mj->mk', lin['Jrp'], dn[:-1] - dn[1:])
            tot += float(np.sum(jd4 * (2 * res[4] + jd4)))
        return -tot


def rel_pose_error(nodes, ref_nodes):
    """Relative pose error used by the parity gate: ||X - Xref||_F / ||Xref||_F over the 7-vector
    storage with canonical quaternion sign, plus max translation / rotation-angle differences."""
    a = np.asarray(nodes, np.float64).copy()
    b = np.asarray(ref_nodes, np.float64).copy()
    a[:, 3:] = lie.quat_canon(a[:, 3:]); b[:, 3:] = lie.quat_canon(b[:, 3:])
    rel = float(np.linalg.norm(a - b) / np.linalg.norm(b))
    dt = float(np.max(np.linalg.norm(a[:, :3] - b[:, :3], axis=1)))
    dq = lie.so3_mul(lie.so3_inv(b[:, 3:]), a[:, 3:])
    ang = float(np.max(np.linalg.norm(lie.so3_log(dq), axis=1)))
    return dict(rel=rel, max_trans=dt, max_rot=ang)
