"""Drop-in mirror of the PyPose helpers of /root/reference/Datasets/transformation.py:72-124 that sit either side of the
PVGO call in train.py (:215-240, :264): same names and semantics, LieTensor shim underneath, and the two Python loops
replaced by device kernels — `motion2pose_pypose` is one prefix-product scan (islam_lie_cumprod), `pose2motion_pypose`
one batched Inv + Mul."""
import torch

from . import pypose_compat as pp


def cvtSE3_pypose(motion):
    """transformation.py:72-86."""
    if isinstance(motion, pp.LieTensor):
        if motion.ltype is pp.SE3_type:
            return motion.clone()
        if motion.ltype is pp.se3_type:
            return motion.Exp()
    else:
        if not isinstance(motion, torch.Tensor):
            motion = torch.tensor(motion)
        if motion.shape[-1] == 6:
            trans = motion[..., :3]
            rot = pp.so3(motion[..., 3:]).Exp().tensor()
            return pp.SE3(torch.cat([trans, rot], dim=-1))
        if motion.shape[-1] == 7:
            return pp.SE3(motion)
    assert False, "Not valid input."


def tartan2kitti_pypose(motion):
    """transformation.py:88-98."""
    motion = cvtSE3_pypose(motion)
    T = [[0., 1., 0., 0.], [0., 0., 1., 0.], [1., 0., 0., 0.], [0., 0., 0., 1.]]
    T = pp.from_matrix(T, ltype=pp.SE3_type).to(motion.device)
    return T @ motion @ T.Inv()


def motion2pose_pypose(motion, T=None):
    """transformation.py:100-113: pose[0] = T, pose[i+1] = pose[i] @ motion[i]  — one scan instead of a Python loop.
    (No autograd through the chain: train.py:221-225 only uses it detached.)"""
    motion = cvtSE3_pypose(motion)
    if T is None:
        T = pp.SE3([0, 0, 0, 0, 0, 0, 1]).to(motion.device)
    else:
        T = cvtSE3_pypose(T).to(motion.device)
    seq = pp.SE3(torch.cat([T.tensor().reshape(1, 7), motion.tensor().detach().reshape(-1, 7)], dim=0))
    return pp.cumprod(seq, dim=0, left=False)


def pose2motion_pypose(pose):
    """transformation.py:115-124: motion[i] = pose[i]^-1 @ pose[i+1]."""
    pose = cvtSE3_pypose(pose)
    return pose[:-1].Inv() @ pose[1:]
