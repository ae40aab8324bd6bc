"""Synthetic Pix3D-shaped batches for the hot path (no dataset in the container).
Field names and shapes follow data/pix3d.py:129-227 of the reference for the fields the render path reads; the CNN
outputs (latent codes, predicted viewpoints) that the out-of-scope encoders would produce are drawn at random."""
import math

import torch

from .options import Options


def _pose_from_angles(az, el, th, scale_dist, cam_dist):
    """world->camera [R|t] built like model/graph.py:273-286: R = Rz Rx Ry P, t = (0, 0, scale_dist * dist)."""
    B = az.shape[0]
    z, o = torch.zeros(B), torch.ones(B)
    Ry = torch.stack([torch.stack([az.cos(), z, az.sin()], -1), torch.stack([z, o, z], -1),
                      torch.stack([-az.sin(), z, az.cos()], -1)], 1)
    Rx = torch.stack([torch.stack([o, z, z], -1), torch.stack([z, el.cos(), -el.sin()], -1),
                      torch.stack([z, el.sin(), el.cos()], -1)], 1)
    Rz = torch.stack([torch.stack([th.cos(), th.sin(), z], -1), torch.stack([-th.sin(), th.cos(), z], -1),
                      torch.stack([z, z, o], -1)], 1)
    P = torch.tensor([[-1., 0, 0], [0, 0, -1], [0, -1, 0]]).expand(B, 3, 3)
    R = Rz @ Rx @ Ry @ P
    t = torch.stack([z, z, scale_dist * cam_dist], -1)
    return torch.cat([R, t[..., None]], -1)


def _targets(ray_idx, H, W, gen):
    """Per-ray targets of a disc-shaped object: rgb U[0,1), mask = disc, normals = unit-sphere normals * mask."""
    B, R = ray_idx.shape
    x = (ray_idx % W).float() + 0.5
    y = torch.div(ray_idx, W, rounding_mode="floor").float() + 0.5
    cx = W / 2 + (torch.rand(B, 1, generator=gen) - 0.5) * W * 0.1
    cy = H / 2 + (torch.rand(B, 1, generator=gen) - 0.5) * H * 0.1
    rad = (0.25 + 0.15 * torch.rand(B, 1, generator=gen)) * min(H, W)
    dx, dy = (x - cx) / rad, (y - cy) / rad
    rr = dx * dx + dy * dy
    mask = (rr < 1).float().unsqueeze(-1)
    nz = torch.sqrt((1 - rr).clamp_min(0))
    normal = torch.stack([dx, dy, -nz], -1) * mask
    rgb = torch.rand(B, R, 3, generator=gen) * mask + (1 - mask)
    return rgb, mask, normal


def make_batch(opt, B, seed=0, pin=True):
    """One host-side training batch (dict of CPU tensors, pinned) with K = opt.data.k_nearest neighbours."""
    gen = torch.Generator().manual_seed(seed)
    H, W = opt.H, opt.W
    R = int(opt.render.rand_sample) if opt.render.rand_sample else H * W
    K = opt.data.k_nearest

    def rays():
        if opt.render.rand_sample:
            return torch.stack([torch.randperm(H * W, generator=gen)[:R] for _ in range(B)])
        return torch.arange(H * W).unsqueeze(0).expand(B, -1).contiguous()

    def camera():
        az = torch.rand(B, generator=gen) * 2 * math.pi
        el = (torch.rand(B, generator=gen) - 0.5) * 0.8
        th = (torch.rand(B, generator=gen) - 0.5) * 0.2
        sd = 1 + 0.1 * (torch.rand(B, generator=gen) - 0.5)
        f = float(opt.camera.focal)
        intr = torch.tensor([[f * W, 0, W / 2], [0, f * H, H / 2], [0, 0, 1.]]).repeat(B, 1, 1)
        return _pose_from_angles(az, el, th, sd, float(opt.camera.dist)), intr, sd

    v = {}
    v["idx"] = torch.arange(B)
    v["ray_idx"] = rays()
    v["rgb_input"], v["mask_input"], v["normal_input"] = _targets(v["ray_idx"], H, W, gen)
    v["pose"], v["intr"], v["scale_dist"] = camera()
    v["proj_latent_sdf"] = torch.randn(B, 64, generator=gen) * 0.3
    v["proj_latent_rgb"] = torch.randn(B, 64, generator=gen) * 0.3
    nn_fields = {k: [] for k in ("ray_idx", "rgb_input", "mask_input", "normal_input", "pose", "intr", "scale_dist",
                                 "proj_latent_rgb")}
    for _ in range(K):
        ri = rays()
        rgb, m, n = _targets(ri, H, W, gen)
        p, i, s = camera()
        for k, t in zip(nn_fields, (ri, rgb, m, n, p, i, s, torch.randn(B, 64, generator=gen) * 0.3)):
            nn_fields[k].append(t)
    for k, lst in nn_fields.items():
        v[k + "_NN"] = torch.stack(lst, dim=-1).contiguous()
    if pin and torch.cuda.is_available():
        v = {k: t.pin_memory() for k, t in v.items()}
    return v


def to_device(batch, device, requires_grad=("pose", "intr", "scale_dist", "proj_latent_sdf", "proj_latent_rgb",
                                            "pose_NN", "intr_NN", "scale_dist_NN", "proj_latent_rgb_NN")):
    """Async host->device copy of a batch; returns (var, bytes copied). CNN-produced fields become grad leaves."""
    var, nbytes = Options(), 0
    for k, t in batch.items():
        d = t.to(device, non_blocking=True)
        nbytes += t.numel() * t.element_size()
        if k in requires_grad:
            d.requires_grad_(True)
        var[k] = d
    return var, nbytes
