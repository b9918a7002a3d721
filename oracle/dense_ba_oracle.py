"""CPU oracle (TEST INFRASTRUCTURE ONLY — imported by tests/, never by the product) for SURVEY.md 8f rank 4:
`scale_from_disp_flow`, /root/reference/dense_ba.py:88-176, the masked one-unknown least squares that turns TartanVO's
up-to-scale translation into metres right before the PVGO back-end (call site TartanVO.py:159-171).

Pinned: tests/golden/scale_golden.npz holds outputs of the reference function itself (run in the build container with a
30-line PyPose stand-in, tests/golden/make_scale_golden.py); tests/test_oracle_scale.py checks this restatement against them.
float64 NumPy, one sample per call, line-by-line citations below."""
import numpy as np


def _qrot(q, p):
    v, w = q[:3], q[3]
    t = 2 * np.cross(v, p)
    return p + w * t + np.cross(v, t)


def scale_from_disp_flow(disp, flow, motion, fx, fy, cx, cy, baseline, depth=None, mask=None, disp_th=1.0):
    """Returns (s, z (H,W), mask (H,W) bool, depth_mask (H,W) bool).  motion: SE3 [t(3), q xyzw] (dense_ba.py:92-95)."""
    disp = np.asarray(disp, np.float64)
    flow = np.asarray(flow, np.float64)
    motion = np.asarray(motion, np.float64)
    H, W = flow.shape[-2:]
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))       # dense_ba.py:98-105
    fu, fv = flow[0] + u, flow[1] + v
    inside = (fu >= 0) & (fu <= W) & (fv >= 0) & (fv <= H)                                     # :66-73 (<= width, inclusive)
    flow_mask = inside & (np.hypot(flow[0], flow[1]) > 0)                                       # :108-109
    m = flow_mask if mask is None else (flow_mask & np.asarray(mask, bool))                     # :110-113
    if depth is None:                                                                           # :115-124
        dmask = ((u - disp) >= 0) & ((u - disp) <= W) & (disp >= disp_th)
        m = m & dmask
        z = np.where(dmask, fx * baseline / np.where(dmask, disp, 1.0), 0.0)
    else:                                                                                       # :126-132
        depth = np.asarray(depth, np.float64)
        dmask = (depth <= fx * baseline) & (depth > 0)
        m = m & dmask
        z = np.where(dmask, depth, 0.0)
    P = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], -1)                                 # z K^-1 [u v 1]   :138-142
    t, q = motion[:3], motion[3:7]
    qi = np.array([-q[0], -q[1], -q[2], q[3]])
    ti = -_qrot(qi, t)                                                                          # T.Inv()          :144-145
    tn = ti / max(np.linalg.norm(ti), 1e-12)                                                    # F.normalize      :146
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    a = K @ tn                                                                                  # :149
    RP = P + qi[3] * 2 * np.cross(qi[:3], P) + np.cross(qi[:3], 2 * np.cross(qi[:3], P))        # R P
    b = RP @ K.T                                                                                # :150
    M1 = a[2] * fu - a[0]; w1 = b[..., 0] - b[..., 2] * fu                                      # :153-156
    M2 = a[2] * fv - a[1]; w2 = b[..., 1] - b[..., 2] * fv
    num = (M1[m] * w1[m]).sum() + (M2[m] * w2[m]).sum()                                         # :159-170
    den = (M1[m] ** 2).sum() + (M2[m] ** 2).sum()
    return num / den, z, m, dmask
