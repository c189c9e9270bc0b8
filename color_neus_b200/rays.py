"""Caller-side ray helpers needed to drive the renderer (synthetic cameras for bench/tests).

Same arithmetic and ray ordering (index = y*W + x) as lib/models/tools/ray_utils.py:7-13 (near_far_from_sphere),
:90-119 (get_rays_at) and lib/utils/transform.py:322-337 (pose_spherical); torch ops on whatever device the
camera tensors live on.  (SURVEY.md section 8f "next" #1: a fused on-device ray generator replaces this.)
"""
import math

import torch


def pose_spherical(theta_deg, phi_deg, radius):
    th, ph = theta_deg / 180.0 * math.pi, phi_deg / 180.0 * math.pi
    t = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], dtype=torch.float32)
    rp = torch.tensor([[1, 0, 0, 0], [0, math.cos(ph), -math.sin(ph), 0], [0, math.sin(ph), math.cos(ph), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    rt = torch.tensor([[math.cos(th), 0, -math.sin(th), 0], [0, 1, 0, 0], [math.sin(th), 0, math.cos(th), 0],
                       [0, 0, 0, 1]], dtype=torch.float32)
    flip = torch.tensor([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=torch.float32)
    return flip @ (rt @ (rp @ t))


def get_rays_at(c2w, focal, H, W, normalize=False, opengl=False):
    """All rays of one pinhole camera -> (rays_o [H,W,3], rays_d [H,W,3])."""
    assert c2w.dim() == 2, "single camera"
    device = c2w.device
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W), torch.linspace(0, H - 1, H), indexing='xy')
    i, j = i.to(device), j.to(device)
    ys, zs = (-1, -1) if opengl else (1, 1)
    dirs = torch.stack([(i - 0.5 * W) / focal[0], ys * (j - 0.5 * H) / focal[1], zs * torch.ones_like(i)], -1)
    if normalize:
        dirs = dirs / torch.norm(dirs, dim=-1).unsqueeze(-1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def near_far_from_sphere(rays_o, rays_d):
    a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return (mid - 1.0).squeeze(-1), (mid + 1.0).squeeze(-1)


def synthetic_camera_rays(H, W, theta_deg=30.0, phi_deg=-30.0, radius=2.8, focal_mul=1.2, device="cpu"):
    """SURVEY.md section 8d synthetic camera: focal = 1.2*W, camera on a sphere looking at the origin,
    NORMALIZE_DIR=True.  Returns flat (rays_o, rays_d, near, far) in y*W+x order."""
    c2w = pose_spherical(theta_deg, phi_deg, radius).to(device)
    focal = torch.tensor([focal_mul * W, focal_mul * W], dtype=torch.float32, device=device)
    o, d = get_rays_at(c2w, focal, H, W, normalize=True)
    o, d = o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()
    near, far = near_far_from_sphere(o, d)
    return o, d, near.contiguous(), far.contiguous()
