"""Ray geometry for the renderer (host side, torch ops on the SELECTED rays only, differentiable).

Mirrors utils/camera.py:157-196 (get_camera_grid + get_center_and_ray, perspective model) and
model/renderer.py:59-68 of the reference, but never builds the H*W ray grid when ray_idx is given.
"""
import torch
import torch.nn.functional as F


def inv3x3(m):
    """Closed-form inverse of [...,3,3] matrices (adjugate / determinant). The reference calls `intr.inverse()`
    (utils/camera.py:166); torch.linalg.inv checks its LU status on the host, i.e. stalls the stream once per render
    and cannot be captured in a CUDA graph. Same values to fp32 rounding; differentiable."""
    r0, r1, r2 = m[..., 0, :], m[..., 1, :], m[..., 2, :]
    c0, c1, c2 = torch.linalg.cross(r1, r2), torch.linalg.cross(r2, r0), torch.linalg.cross(r0, r1)
    det = (r0 * c0).sum(-1, keepdim=True)
    return torch.stack([c0 / det, c1 / det, c2 / det], dim=-1)


def pixel_rays(pose, intr, H, W, ray_idx=None):
    """pose [B,3,4] world->camera, intr [B,3,3]; ray_idx [B,R] int64 flat pixel ids (row-major) or None = all.
    -> cam_loc [B,3], unit ray_dirs [B,R,3], depth_fac [B,R] (ray length -> depth factor)."""
    B = pose.shape[0]
    dev, dt = pose.device, pose.dtype
    if ray_idx is None:
        idx = torch.arange(H * W, device=dev).unsqueeze(0).expand(B, -1)
    else:
        idx = ray_idx
    px = (idx % W).to(dt) + 0.5
    py = torch.div(idx, W, rounding_mode="floor").to(dt) + 0.5
    pix = torch.stack([px, py, torch.ones_like(px)], dim=-1)              # [B,R,3]
    cam = pix @ inv3x3(intr).transpose(-1, -2)
    Rm, t = pose[..., :3], pose[..., 3]
    center = -(Rm.transpose(-1, -2) @ t.unsqueeze(-1))[..., 0]            # camera centre in world coords
    world = cam @ Rm + center.unsqueeze(1)                               # inverse pose applied to the pixel points
    raw = world - center.unsqueeze(1)
    dirs = F.normalize(raw, dim=-1)
    depth_fac = dirs.norm(dim=-1) / raw.norm(dim=-1)
    return center, dirs, depth_fac
