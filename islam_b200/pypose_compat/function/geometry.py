"""pypose.function.geometry names imported at /root/reference/dense_ba.py:5 (used only by the sparse reprojection class
that train.py never enables).  Plain torch, off the hot path."""
import torch


def point2pixel(points, intrinsics, extrinsics=None):
    if extrinsics is not None:
        points = extrinsics.unsqueeze(-2) @ points
    uv = points[..., :2] / points[..., 2:3]
    fx, fy = intrinsics[..., 0, 0], intrinsics[..., 1, 1]
    cx, cy = intrinsics[..., 0, 2], intrinsics[..., 1, 2]
    f = torch.stack([fx, fy], -1).unsqueeze(-2)
    c = torch.stack([cx, cy], -1).unsqueeze(-2)
    return uv * f + c


def reprojerr(points, pixels, intrinsics, extrinsics=None, reduction='none'):
    err = point2pixel(points, intrinsics, extrinsics) - pixels
    if reduction == 'norm':
        return err.norm(dim=-1)
    return err
