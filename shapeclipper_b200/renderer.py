"""Renderer with the reference's constructor/forward signature (model/renderer.py:40-185), evaluated by the fused
sm_100a render kernels. Drop-in for `model.renderer` (see shapeclipper_b200.shim).

What runs where:
  * ray geometry from pose/intrinsics for the selected pixels: torch ops on [B,R,3] (camera.py), differentiable;
  * stratified depths, sample points, posenc, SDF MLP + its spatial gradient, Laplace density, RGB MLP, alpha
    compositing and the rgb/mask/depth/normal reductions, forward AND backward: one CUDA kernel each way;
  * eikonal samples: the SDF point-query kernel (get_conditional_output) + a norm.
The CPU-generator draws happen in the reference's order (rand -> randint -> uniform_) so equal seeds give equal
samples.
"""
import torch
import torch.nn as nn

from . import camera
from .implicit import LaplaceDensity
from .render_fn import render_rays


class UniformSampler(nn.Module):
    """Depth sampler of the reference (model/renderer.py:8-37). The depths themselves are produced inside the render
    kernel; this class keeps the attribute surface and restates z for the eikonal sample (one depth per ray)."""

    def __init__(self, opt):
        super().__init__()
        self.N_samples = opt.render.n_samples_uniform

    @staticmethod
    def depth_at(opt, scale_dist_per_ray, t_vals, idx, u_at_idx):
        """z of sample `idx` [N] for rays with scale_dist_per_ray [N]; u_at_idx = jitter of that sample or None."""
        c = opt.camera.dist * scale_dist_per_ray
        near, far = c - 0.7, c + 0.7
        S = t_vals.shape[0]

        def zb(i):
            t = t_vals[i.clamp(0, S - 1)]
            return near * (1.0 - t) + far * t
        z = zb(idx)
        if u_at_idx is None:
            return z
        upper = torch.where(idx < S - 1, 0.5 * (zb(idx + 1) + z), z)
        lower = torch.where(idx > 0, 0.5 * (z + zb(idx - 1)), z)
        return lower + (upper - lower) * u_at_idx


class Renderer(nn.Module):
    def __init__(self, opt, sdf_network, rgb_network):
        super().__init__()
        self.bg_color = opt.data.bgcolor
        self.eik_range = opt.arch.impl_sdf.eikonal_sample_range
        self.normal_model = opt.render.normal_model
        self.sdf_network = sdf_network
        self.rgb_network = rgb_network
        self.density = LaplaceDensity(params_init={'beta': opt.arch.impl_sdf.beta_init})
        if opt.render.sampler != 'uniform':
            raise NotImplementedError
        if self.normal_model != 'volume':
            raise NotImplementedError("only render.normal_model == 'volume' (the reference default) is implemented")
        self.ray_sampler = UniformSampler(opt)
        self.N_samples = opt.render.n_samples_uniform
        self._t_cache = {}

    def _t_vals(self, S, dev):
        key = (S, str(dev))
        if self._t_cache.get("key") != key:
            self._t_cache = {"key": key, "t": torch.linspace(0., 1., steps=S).to(dev)}
        return self._t_cache["t"]

    def forward(self, opt, pose, intr, scale_dist, proj_latent_sdf, proj_latent_rgb, ray_idx=None, training=True,
                visualize=False, eikonal=True):
        """The reference's signature plus `eikonal` (default True = reference behaviour): a caller that discards
        grad_eikonal — the reference's neighbour-view render, model/graph.py:207 — can pass False to skip the eikonal point
        queries; the generator draws still happen, so the random stream stays in step with the reference."""
        if opt.camera.model != "perspective":
            raise NotImplementedError("only the perspective camera of the reference config is implemented")
        dev = pose.device
        S = self.N_samples
        cam_loc, ray_dirs, depth_fac = camera.pixel_rays(pose, intr, opt.H, opt.W, ray_idx)
        B, R = ray_dirs.shape[0], ray_dirs.shape[1]

        # CPU-generator draws, reference order (renderer.py:29,33,158). `opt.render.device_rng` draws the same
        # distributions from the CUDA generator instead: no host work / pageable copy in the step, and the step becomes
        # capturable in a CUDA graph (throughput mode; seeds then no longer reproduce the reference's samples).
        rng_dev = dev if getattr(opt.render, "device_rng", False) else "cpu"
        u = torch.rand(B * R, S, device=rng_dev).to(dev) if training else None
        eik_idx = torch.randint(S, (B * R,), device=rng_dev).to(dev)
        t_vals = self._t_vals(S, dev)

        cfg = dict(n_samples=S, beta_min=float(self.density.beta_min), cam_dist=float(opt.camera.dist), half_range=0.7,
                   bg_color=float(self.bg_color), normal_pow=float(opt.reg.normal_pow))
        rgb, mask, mask_hard, depth, normal = render_rays(
            cfg, self.density.beta, cam_loc, ray_dirs, depth_fac, scale_dist, proj_latent_sdf, proj_latent_rgb,
            t_vals, u, self.sdf_network, self.rgb_network)

        grad_eikonal = None
        if training:
            uni = torch.empty(B * R, 3, device=rng_dev).uniform_(self.eik_range[0], self.eik_range[1]).to(dev).reshape(B, R, 3)
        if training and eikonal:
            eik_pts = camera.eikonal_points(cam_loc, ray_dirs, scale_dist, t_vals, u, eik_idx, uni, opt.camera.dist)
            _, _, g = self.sdf_network.get_conditional_output(opt, B, eik_pts, proj_latent_sdf, compute_grad=True)
            grad_eikonal = g.norm(2, dim=1)
        if visualize:
            return (rgb, mask, mask_hard, depth, normal, grad_eikonal) + self._visualize_rays(
                opt, cam_loc, ray_dirs, scale_dist, t_vals, u, proj_latent_sdf, proj_latent_rgb)
        return rgb, mask, mask_hard, depth, normal, grad_eikonal

    @torch.no_grad()
    def _visualize_rays(self, opt, cam_loc, ray_dirs, scale_dist, t_vals, u, z_sdf, z_rgb, n_vis=200):
        """Per-sample points / opacity / colour of 200 random rays (model/renderer.py:174-183, 212-215): debug output of
        evaluate(visualize=True), not the hot path. SDF + features come from the point-query kernel; the per-sample
        colours are the only place the standalone RGBNetwork.forward runs."""
        B, R, S = ray_dirs.shape[0], ray_dirs.shape[1], self.N_samples
        dev = ray_dirs.device
        idx = torch.randperm(R)[:n_vis].to(dev)                       # CPU generator, drawn last as in the reference
        n = idx.numel()
        d = ray_dirs[:, idx]                                           # [B,n,3]
        s_idx = torch.arange(S, device=dev).repeat(B * n)
        sd = scale_dist.reshape(B, 1, 1).expand(B, n, S).reshape(-1)
        u_sel = None
        if u is not None:
            u_sel = u.reshape(B, R, S)[:, idx].reshape(-1)
        z = UniformSampler.depth_at(opt, sd, t_vals, s_idx, u_sel).reshape(B, n, S)
        pts = cam_loc.reshape(B, 1, 1, 3) + z.unsqueeze(-1) * d.unsqueeze(2)        # [B,n,S,3]
        flat = pts.reshape(-1, 3).contiguous()
        sdf, feat, _ = self.sdf_network.get_conditional_output(opt, B, flat, z_sdf, compute_grad=False)
        sigma = self.density(sdf).reshape(B, n, S)
        delta = torch.cat([z[..., 1:] - z[..., :-1], torch.zeros_like(z[..., :1])], dim=-1)
        opacity = (1 - torch.exp(-delta * sigma)).reshape(B, n * S, 1)
        lat = z_rgb.reshape(B, 1, -1).expand(B, n * S, z_rgb.shape[-1]).reshape(B * n * S, -1)
        rgb = self.rgb_network(flat, lat, feat).reshape(B, n * S, 3)
        transparency = torch.cat([opacity, 1 - opacity, torch.zeros_like(opacity)], dim=-1)
        return pts.reshape(B, n * S, 3), transparency, torch.cat([rgb, opacity], dim=-1)
