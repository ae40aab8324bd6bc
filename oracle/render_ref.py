"""ORACLE (test infrastructure, not product code): CPU restatement of ShapeClipper's render path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product (shapeclipper_b200/) never does and fails loudly without its CUDA library.

What is restated (reference file:line, relative to /root/reference):
  posenc                model/implicit.py:7-52        (Embedder / get_embedder)
  sdf_mlp               model/implicit.py:138-161     (SDFNetwork.forward)
  sdf_query             model/implicit.py:163-189     (SDFNetwork.get_conditional_output)
  rgb_mlp               model/implicit.py:220-239     (RGBNetwork.forward)
  laplace_density       model/implicit.py:65-83       (LaplaceDensity)
  camera_rays           utils/camera.py:157-196       (get_camera_grid + get_center_and_ray, perspective)
  depth_samples         model/renderer.py:13-37       (UniformSampler.get_z_vals)
  composite             model/renderer.py:187-209     (Renderer.volume_rendering)
  render                model/renderer.py:57-185      (Renderer.forward)
  draw_render_rng       model/renderer.py:29,33,158   (order of the CPU-generator draws)
  level_grid            utils/eval_3D.py:9-38         (get_dense_3D_grid + compute_level_grid)
  normalize_pc, fscore  utils/eval_3D.py:40-49,105-121

Parity pinning: the reference has no tests or golden vectors. This restatement is
pinned against the reference itself, imported unmodified in the build container
(tests/test_oracle_render.py) and against fixtures that import produced
(tests/golden/*.pt, generator tests/gen_golden.py).

Everything is dtype-generic: run it in float32 to mirror the reference, in float64 for a truth run.
"""
import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class RenderCfg:
    """The option leaves the render path reads (options/pix3d/config.yaml)."""
    n_samples: int = 64            # render.n_samples_uniform
    cam_dist: float = 5.0          # camera.dist
    depth_half_range: float = 0.7  # hard-coded at model/renderer.py:16-17
    bg_color: float = 1.0          # data.bgcolor
    normal_pow: float = 1.0        # reg.normal_pow
    eik_lo: float = -1.0           # arch.impl_sdf.eikonal_sample_range
    eik_hi: float = 1.0
    n_freq: int = 6                # arch.impl_*.pos_enc
    symmetry: bool = True          # arch.force_symmetry
    skip_in: tuple = (1, 2)        # arch.impl_sdf.skip_connection
    beta_min: float = 1e-4         # LaplaceDensity(beta_min)


# ----------------------------------------------------------------------------- MLPs

def posenc(x, n_freq=6):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]  -> [..., 3 + 6 L]."""
    parts = [x]
    for k in range(n_freq):
        f = float(2 ** k)
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, dim=-1)


def _mirror_x(pts):
    return torch.cat([pts[..., :1].abs(), pts[..., 1:]], dim=-1)


def sdf_mlp(params, pts, latent, cfg=RenderCfg()):
    """params: {'lin{l}.weight','lin{l}.bias'} l=0..5; pts [N,3]; latent [N,64] -> [N,65]."""
    n_lin = len([k for k in params if k.endswith(".weight")])
    p = _mirror_x(pts) if cfg.symmetry else pts
    net_in = torch.cat([posenc(p, cfg.n_freq), latent], dim=-1)
    h = net_in
    for l in range(n_lin):
        if l in cfg.skip_in:
            h = torch.cat([h, net_in], dim=-1) / math.sqrt(2)
        h = F.linear(h, params["lin%d.weight" % l], params["lin%d.bias" % l])
        if l < n_lin - 1:
            h = F.softplus(h, beta=100, threshold=20)
    return h


def rgb_mlp(params, pts, latent, feat, cfg=RenderCfg()):
    """pts [N,3]; latent [N,64]; feat [N,64] -> sigmoid rgb [N,3]."""
    n_lin = len([k for k in params if k.endswith(".weight")])
    p = _mirror_x(pts) if cfg.symmetry else pts
    h = torch.cat([posenc(p, cfg.n_freq), latent, feat], dim=-1)
    for l in range(n_lin):
        h = F.linear(h, params["lin%d.weight" % l], params["lin%d.bias" % l])
        if l < n_lin - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)


def effective_beta(beta_param, cfg=RenderCfg()):
    return beta_param.abs() + cfg.beta_min


def laplace_density(sdf, beta):
    """sigma(s) = 1/beta * (0.5 e^{-s/beta} if s >= 0 else 1 - 0.5 e^{s/beta})."""
    e = 0.5 * torch.exp(-sdf.abs() / beta)
    return torch.where(sdf >= 0, e, 1 - e) / beta


def sdf_query(params, pts_flat, latent, batch_size, want_grad=True, cfg=RenderCfg()):
    """pts_flat [B*N,3] batch-major, latent [B,64] -> (sdf [B*N,1], feat [B*N,64], d sdf/d pts or None).

    With want_grad the latent is detached, as the reference does (model/implicit.py:168-169)."""
    n = pts_flat.shape[0] // batch_size
    lat = latent.unsqueeze(1).expand(batch_size, n, latent.shape[-1]).reshape(batch_size * n, -1)
    if want_grad:
        lat = lat.detach()
        if not pts_flat.requires_grad:
            pts_flat = pts_flat.detach().requires_grad_(True)
    out = sdf_mlp(params, pts_flat, lat, cfg)
    sdf, feat = out[:, :1], out[:, 1:]
    grad = None
    if want_grad:
        grad = torch.autograd.grad(sdf, pts_flat, torch.ones_like(sdf), create_graph=True, retain_graph=True)[0]
    return sdf, feat, grad


# ----------------------------------------------------------------------------- camera

def camera_rays(pose, intr, H, W, ray_idx=None):
    """Perspective rays. pose [B,3,4] world->cam, intr [B,3,3].
    -> origin [B,3], unit dirs [B,R,3], depth_fac [B,R] (= 1/|unnormalised ray|).
    Mirrors the reference's operation order: unproject every pixel centre, map through the inverse
    pose, subtract the mapped camera centre, then (optionally) gather ray_idx."""
    B = pose.shape[0]
    dt, dev = pose.dtype, pose.device
    ys = torch.arange(H, dtype=dt, device=dev) + 0.5
    xs = torch.arange(W, dtype=dt, device=dev) + 0.5
    Y, X = torch.meshgrid(ys, xs, indexing="ij")
    pix = torch.stack([X, Y, torch.ones_like(X)], dim=-1).view(1, H * W, 3).expand(B, -1, -1)
    cam_pts = pix @ torch.linalg.inv(intr).transpose(-1, -2)
    R, t = pose[..., :3], pose[..., 3:]
    R_inv = R.transpose(-1, -2)
    t_inv = (-R_inv @ t)[..., 0]                                  # [B,3] camera centre in world
    inv_T = torch.cat([R_inv, t_inv[..., None]], dim=-1).transpose(-1, -2)   # [B,4,3]
    hom = torch.cat([cam_pts, torch.ones_like(cam_pts[..., :1])], dim=-1)
    world = hom @ inv_T
    zero_h = torch.cat([torch.zeros(B, 1, 3, dtype=dt, device=dev), torch.ones(B, 1, 1, dtype=dt, device=dev)], -1)
    origin = zero_h @ inv_T                                       # [B,1,3]
    raw = world - origin
    dirs = F.normalize(raw, dim=-1)
    depth_fac = dirs.norm(dim=-1) / raw.norm(dim=-1)
    if ray_idx is not None:
        dirs = dirs.gather(1, ray_idx[..., None].expand(-1, -1, 3))
        depth_fac = depth_fac.gather(1, ray_idx)
    return origin[:, 0], dirs, depth_fac


# ----------------------------------------------------------------------------- sampling

def draw_render_rng(n_rays_total, n_samples, training, cfg=RenderCfg(), generator=None):
    """The CPU-generator draws of one Renderer.forward, in the reference's order:
    rand([BR,S]) (training only) -> randint(S,[BR]) (always) -> uniform_(lo,hi)[BR,3] (training only)."""
    u = torch.rand(n_rays_total, n_samples, generator=generator) if training else None
    eik_idx = torch.randint(n_samples, (n_rays_total,), generator=generator)
    eik_pts = None
    if training:
        eik_pts = torch.empty(n_rays_total, 3).uniform_(cfg.eik_lo, cfg.eik_hi, generator=generator)
    return u, eik_idx, eik_pts


def depth_samples(scale_dist, n_rays, cfg=RenderCfg(), u=None):
    """scale_dist [B] -> z [B*n_rays, S]; u = stratified jitter in [0,1) or None (eval: bin edges)."""
    c = (cfg.cam_dist * scale_dist).repeat_interleave(n_rays).unsqueeze(-1)
    near, far = c - cfg.depth_half_range, c + cfg.depth_half_range
    t = torch.linspace(0.0, 1.0, cfg.n_samples, dtype=torch.float32).to(scale_dist.device)
    z = near * (1.0 - t) + far * t
    if u is not None:
        mids = 0.5 * (z[:, 1:] + z[:, :-1])
        upper = torch.cat([mids, z[:, -1:]], dim=-1)
        lower = torch.cat([z[:, :1], mids], dim=-1)
        z = lower + (upper - lower) * u.to(z.dtype)
    return z


def composite(z, sigma):
    """z [N,S], sigma [N,S] -> weights, alpha. delta_{S-1} = 0; T_i = exp(-sum_{j<i} delta_j sigma_j)."""
    delta = torch.cat([z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])], dim=-1)
    energy = delta * sigma
    before = torch.cat([torch.zeros_like(energy[:, :1]), energy[:, :-1]], dim=-1)
    alpha = 1 - torch.exp(-energy)
    trans = torch.exp(-torch.cumsum(before, dim=-1))
    return alpha * trans, alpha


# ----------------------------------------------------------------------------- the renderer

def render(sdf_params, rgb_params, beta_param, pose, intr, scale_dist, z_sdf, z_rgb, H, W,
           ray_idx=None, training=True, rng=None, cfg=RenderCfg()):
    """One Renderer.forward. rng = (u, eik_idx, eik_pts) from draw_render_rng (drawn here if None).
    Returns dict(rgb[B,R,3], mask[B,R,1], mask_hard[B,R,1], depth[B,R,1], normal[B,R,3],
                 grad_eik[B*2R] or None, + z, weights, sdf for debugging)."""
    B = pose.shape[0]
    S = cfg.n_samples
    origin, dirs, depth_fac = camera_rays(pose, intr, H, W, ray_idx)
    R = dirs.shape[1]
    if rng is None:
        rng = draw_render_rng(B * R, S, training, cfg)
    u, eik_idx, eik_pts = rng
    dev = pose.device
    o = origin.unsqueeze(1).expand(B, R, 3).reshape(-1, 3)
    d = dirs.reshape(-1, 3)
    z = depth_samples(scale_dist, R, cfg, u.to(dev) if (training and u is not None) else None)
    pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)

    beta = effective_beta(beta_param, cfg)
    with torch.enable_grad():
        if not pts.requires_grad:
            pts.requires_grad_(True)
        sdf, feat, _ = sdf_query(sdf_params, pts, z_sdf, B, want_grad=False, cfg=cfg)
        sigma = laplace_density(sdf, beta)
        n_flat = -torch.autograd.grad(sigma, pts, torch.ones_like(sigma), create_graph=True, retain_graph=True)[0]
    lat_rgb = z_rgb.unsqueeze(1).expand(B, R * S, z_rgb.shape[-1]).reshape(B * R * S, -1)
    color = rgb_mlp(rgb_params, pts, lat_rgb, feat, cfg).reshape(-1, S, 3)

    w, alpha = composite(z, laplace_density(sdf, beta).reshape(-1, S))
    depth = (w * (z * depth_fac.reshape(-1, 1))).sum(1)
    n_s = F.normalize(n_flat, dim=-1).reshape(-1, S, 3)
    normal = F.normalize(((w.unsqueeze(-1) ** cfg.normal_pow) * n_s).sum(1), dim=-1)
    acc = w.sum(-1)
    rgb = (w.unsqueeze(-1) * color).sum(1) + (1.0 - acc).unsqueeze(1) * cfg.bg_color

    grad_eik = None
    if training:
        z_eik = z.gather(1, eik_idx.to(dev).unsqueeze(-1))
        near_pts = (o + z_eik * d).reshape(B, R, 3)
        uni = eik_pts.to(dev).to(z.dtype).reshape(B, R, 3)
        e_pts = torch.cat([uni, near_pts], dim=1).reshape(-1, 3)
        _, _, g = sdf_query(sdf_params, e_pts, z_sdf, B, want_grad=True, cfg=cfg)
        grad_eik = g.norm(2, dim=1)

    return dict(rgb=rgb.view(B, R, 3), mask=acc.view(B, R, 1), mask_hard=(acc > 0.5).to(acc.dtype).view(B, R, 1),
                depth=depth.view(B, R, 1), normal=normal.view(B, R, 3), grad_eik=grad_eik,
                z=z, weights=w, alpha=alpha, sdf=sdf, color=color, n_flat=n_flat, pts=pts)


# ----------------------------------------------------------------------------- evaluation helpers

@torch.no_grad()
def level_grid(sdf_params, z_sdf, vox_res, lo=-0.6, hi=0.6, cfg=RenderCfg()):
    """SDF on the (N+1)^3 lattice over [lo,hi]^3, index order [B, ix, iy, iz]."""
    B = z_sdf.shape[0]
    g = torch.linspace(lo, hi, vox_res + 1, device=z_sdf.device)
    n = vox_res + 1
    out = []
    for i in range(n):
        X, Y, Z = torch.meshgrid(g[i:i + 1], g, g, indexing="ij")
        pts = torch.stack([X, Y, Z], dim=-1).reshape(1, -1, 3).expand(B, -1, -1).reshape(-1, 3).to(z_sdf.dtype)
        lat = z_sdf.unsqueeze(1).expand(B, n * n, -1).reshape(B * n * n, -1)
        out.append(sdf_mlp(sdf_params, pts, lat, cfg)[:, 0].view(B, 1, n, n))
    return torch.cat(out, dim=1)


def normalize_pc(pc):
    c = pc - pc.mean(dim=1, keepdim=True)
    ext = torch.stack([c[:, :, 0].amax(-1) - c[:, :, 0].amin(-1), c[:, :, 1].amax(-1) - c[:, :, 1].amin(-1)], -1)
    return c / (ext.amax(-1)[:, None, None] + 1.0e-7)


def fscore(dist1, dist2, thresholds=(0.005, 0.01, 0.02, 0.05, 0.1, 0.2)):
    cols = []
    for th in thresholds:
        p = (dist1 < th).float().mean(1)
        r = (dist2 < th).float().mean(1)
        f = 2 * p * r / (p + r)
        cols.append(torch.where(torch.isnan(f), torch.zeros_like(f), f))
    return torch.stack(cols, dim=1)
