"""Camera label construction for the 25-float conditioning vector c = [cam2world 4x4 | intrinsics 3x3].

Host-side input construction only (callers of the hot path); follows camera_utils.py:68-86
(LookAtPoseSampler.sample with zero stddev), :118-137 (create_cam2world_matrix) and :140-149
(FOV_to_intrinsics, which uses pi ~ 3.14159 and sqrt(2) ~ 1.414 literally).
"""
import math

import torch


def _normalize(v):
    return v / torch.linalg.norm(v, dim=-1, keepdim=True)


def look_at_pose(horizontal, vertical, lookat, radius):
    """-> cam2world [4,4] float32 for a camera on a sphere of `radius` looking at `lookat`."""
    h = torch.tensor(float(horizontal), dtype=torch.float32)
    v = torch.tensor(float(vertical), dtype=torch.float32).clamp(1e-5, math.pi - 1e-5)
    theta = h
    phi = torch.arccos(1 - 2 * (v / math.pi))
    origin = torch.stack([radius * torch.sin(phi) * torch.cos(math.pi - theta),
                          radius * torch.cos(phi),
                          radius * torch.sin(phi) * torch.sin(math.pi - theta)])
    fwd = _normalize(_normalize(torch.as_tensor(lookat, dtype=torch.float32) - origin))   # normalised twice, like the reference (:85 and :124)
    up = torch.tensor([0., 1., 0.])
    right = -_normalize(torch.linalg.cross(up, fwd))
    up = _normalize(torch.linalg.cross(fwd, right))
    rot = torch.eye(4)
    rot[:3, :3] = torch.stack([right, up, fwd], dim=-1)
    trans = torch.eye(4)
    trans[:3, 3] = origin
    return trans @ rot


def fov_to_intrinsics(fov_degrees):
    focal = float(1 / (math.tan(fov_degrees * 3.14159 / 360) * 1.414))
    return torch.tensor([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], dtype=torch.float32)
