"""SDFNetwork / RGBNetwork / LaplaceDensity with the reference's constructor + forward signatures, parameter
names and state_dict keys (model/implicit.py:65-239), evaluated by the fused sm_100a kernels.

Weight layout is nn.Linear's ([out,in] fp32), so reference checkpoints load unchanged
(`sdf_network.lin{0..5}.{weight,bias}`, `rgb_network.lin{0..3}.*`, `renderer.density.beta`).
The kernels are specialised to the architecture of options/pix3d/config.yaml; anything else raises.
"""
import math

import torch
import torch.nn as nn

from . import _render_native as rn
from . import _lib
from .render_fn import sdf_query


def _require(cond, what):
    if not cond:
        raise NotImplementedError("shapeclipper_b200 kernels are specialised to the reference config; unsupported: " + what)


class Density(nn.Module):
    def __init__(self, params_init={}):
        super().__init__()
        for name, value in params_init.items():
            setattr(self, name, nn.Parameter(torch.tensor(value)))

    def forward(self, sdf, beta=None):
        return self.density_func(sdf, beta=beta)


class LaplaceDensity(Density):
    """sigma(s) = (1/beta) * (0.5 exp(-s/beta) if s >= 0 else 1 - 0.5 exp(s/beta)); beta = |beta_param| + beta_min.
    Inside Renderer.forward the density is fused into the render kernel; this torch form serves direct callers."""

    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__(params_init=params_init)
        self.beta_min = torch.tensor(beta_min)

    def density_func(self, sdf, beta=None):
        if beta is None:
            beta = self.get_beta()
        half = 0.5 * torch.exp(-sdf.abs() / beta)
        return torch.where(sdf >= 0, half, 1 - half) / beta

    def get_beta(self):
        return self.beta.abs() + self.beta_min.to(self.beta.device)


class SDFNetwork(nn.Module):
    def __init__(self, opt):
        super().__init__()
        cfg = opt.arch.impl_sdf
        self.force_symmetry = opt.arch.force_symmetry
        self.proj_latent_dim = cfg.proj_latent_dim
        self.n_hidden = cfg.n_hidden_layers
        self.n_channel = cfg.n_channels
        self.skip_in = list(cfg.skip_connection)
        _require(self.force_symmetry is True, "arch.force_symmetry != true")
        _require(self.proj_latent_dim == 64 and self.n_channel == 64 and self.n_hidden == 5, "impl_sdf sizes != 64/64/5")
        _require(cfg.pos_enc == 6 and self.skip_in == [1, 2] and not cfg.weight_norm, "impl_sdf pos_enc/skip/weight_norm")
        pe_dim = 3 + 3 * 2 * cfg.pos_enc
        d_in = pe_dim + self.proj_latent_dim
        dims = [d_in] + [self.n_channel] * self.n_hidden + [1 + self.n_channel]
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            fan_in = dims[l] + (d_in if l in self.skip_in else 0)
            fan_out = dims[l + 1]
            lin = nn.Linear(fan_in, fan_out)
            if cfg.geometric_init:
                # sphere initialisation of SAL/IGR as used by VolSDF (reference: model/implicit.py:114-128);
                # the order of the in-place draws matches the reference so equal seeds give equal weights
                std = math.sqrt(2) / math.sqrt(fan_out)
                if l == self.num_layers - 2:
                    nn.init.normal_(lin.weight, mean=math.sqrt(math.pi) / math.sqrt(fan_in), std=0.0001)
                    nn.init.constant_(lin.bias, -cfg.init_sphere_radius)
                elif l == 0:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.constant_(lin.weight[:, 3:], 0.0)
                    nn.init.normal_(lin.weight[:, :3], 0.0, std)
                elif l in self.skip_in:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, std)
                    nn.init.constant_(lin.weight[:, -(d_in - 3):], 0.0)
                else:
                    nn.init.constant_(lin.bias, 0.0)
                    nn.init.normal_(lin.weight, 0.0, std)
            setattr(self, "lin%d" % l, lin)
        self.softplus = nn.Softplus(beta=100)   # kept for state/attribute parity; evaluated in-kernel

    def linears(self):
        return [getattr(self, "lin%d" % l) for l in range(self.num_layers - 1)]

    def forward(self, points_raw, proj_latent):
        """points_raw [N,3], proj_latent [N,64] (one latent row per point) -> [N,65]."""
        N = points_raw.shape[0]
        with torch.no_grad():
            change = (proj_latent[1:] != proj_latent[:-1]).any(dim=1)
            starts = torch.cat([torch.zeros(1, dtype=torch.long, device=change.device),
                                change.nonzero().flatten() + 1])
        n_runs = starts.numel()
        if N % n_runs == 0 and bool((starts == torch.arange(n_runs, device=starts.device) * (N // n_runs)).all()):
            lat = proj_latent[starts]
            sdf, feat, _ = sdf_query(self, None, n_runs, points_raw, lat, want_grad=False, detach_latent=False)
        else:   # arbitrary per-point latents: every point is its own "image"
            sdf, feat, _ = sdf_query(self, None, N, points_raw, proj_latent, want_grad=False, detach_latent=False)
        return torch.cat([sdf, feat], dim=-1)

    def get_conditional_output(self, opt, batch_size, points_flat, proj_latent, compute_grad=True):
        """points_flat [B*N,3] batch-major, proj_latent [B,64] -> (sdf [B*N,1], feat [B*N,64], d sdf/d x or None).
        As in the reference the latent is detached when compute_grad is set."""
        assert proj_latent.shape[1] == opt.arch.impl_sdf.proj_latent_dim
        return sdf_query(self, opt, batch_size, points_flat, proj_latent, want_grad=bool(compute_grad),
                         detach_latent=bool(compute_grad))


class RGBNetwork(nn.Module):
    def __init__(self, opt):
        super().__init__()
        cfg = opt.arch.impl_rgb
        self.force_symmetry = opt.arch.force_symmetry
        self.proj_latent_dim = cfg.proj_latent_dim
        self.n_hidden = cfg.n_hidden_layers
        self.n_sdf_channel = opt.arch.impl_sdf.n_channels
        self.n_channel = cfg.n_channels
        _require(self.proj_latent_dim == 64 and self.n_channel == 64 and self.n_hidden == 3 and self.n_sdf_channel == 64,
                 "impl_rgb sizes != 64/64/3")
        _require(cfg.pos_enc == 6 and not cfg.weight_norm, "impl_rgb pos_enc/weight_norm")
        d_in = 3 + 3 * 2 * cfg.pos_enc + self.proj_latent_dim + self.n_sdf_channel
        dims = [d_in] + [self.n_channel] * self.n_hidden + [3]
        self.num_layers = len(dims)
        for l in range(self.num_layers - 1):
            setattr(self, "lin%d" % l, nn.Linear(dims[l], dims[l + 1]))
        self.relu = nn.ReLU()
        self.sigmoid = nn.Sigmoid()

    def linears(self):
        return [getattr(self, "lin%d" % l) for l in range(self.num_layers - 1)]

    def forward(self, points_raw, proj_latent, sdf_feature):
        """points_raw [N,3], proj_latent [N,64], sdf_feature [N,64] -> rgb [N,3] (model/implicit.py:220-239).
        NOT the hot path: inside Renderer.forward the RGB MLP is fused into the render kernel. This standalone form only
        serves the 200-ray debug output of `visualize=True` (model/renderer.py:174-183) and direct callers; it is a few
        torch ops on the caller's device."""
        x = torch.cat([points_raw[:, :1].abs(), points_raw[:, 1:]], dim=-1) if self.force_symmetry else points_raw
        h = torch.cat([positional_encoding(x, 6), proj_latent, sdf_feature], dim=-1)
        lins = self.linears()
        for l, lin in enumerate(lins):
            h = lin(h)
            if l < len(lins) - 1:
                h = self.relu(h)
        return self.sigmoid(h)


def positional_encoding(x, n_freq):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)] (model/implicit.py:28-38), 3 -> 3 + 6 L."""
    out = [x]
    for f in range(n_freq):
        out += [torch.sin(x * float(2 ** f)), torch.cos(x * float(2 ** f))]
    return torch.cat(out, dim=-1)
