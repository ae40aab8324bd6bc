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
    """Ray geometry of the selected pixels: CUDA tensors go through the fused kernels (one launch forward, two backward —
    sc_pixel_rays_forward / _backward), anything else through the torch restatement below (same values to fp32 rounding)."""
    if pose.is_cuda:
        return _PixelRays.apply(pose, intr, ray_idx, int(H), int(W))
    return pixel_rays_torch(pose, intr, H, W, ray_idx)


class _PixelRays(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pose, intr, ray_idx, H, W):
        from . import _lib, _render_native as rn
        L = _lib.lib()
        _lib.require_cuda(pose, intr, ray_idx)
        B = pose.shape[0]
        R = ray_idx.shape[1] if ray_idx is not None else H * W
        p, k = rn._f32c(pose.detach()), rn._f32c(intr.detach())
        idx = ray_idx.contiguous() if ray_idx is not None else None
        if idx is not None and idx.dtype != torch.int64:
            idx = idx.long()
        dev = pose.device
        center = torch.empty(B, 3, device=dev); dirs = torch.empty(B, R, 3, device=dev); fac = torch.empty(B, R, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.sc_pixel_rays_forward(_lib.ptr(p), _lib.ptr(k), _lib.ptr(idx), B, R, W, _lib.ptr(center), _lib.ptr(dirs),
                                               _lib.ptr(fac), _lib.stream_of(p)), "sc_pixel_rays_forward")
        rn.TIMERS.count()
        ctx.save_for_backward(p, k, idx if idx is not None else p)
        ctx.meta = (B, R, W, idx is not None)
        ctx.set_materialize_grads(False)
        return center, dirs, fac

    @staticmethod
    def backward(ctx, g_center, g_dirs, g_fac):
        from . import _lib, _render_native as rn
        L = _lib.lib()
        p, k, idx = ctx.saved_tensors
        B, R, W, has_idx = ctx.meta
        dev = p.device
        c = lambda t: rn._f32c(t) if t is not None else None
        g_center, g_dirs, g_fac = c(g_center), c(g_dirs), c(g_fac)
        need_pose, need_intr = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        pose_bar = torch.empty(B, 3, 4, device=dev) if need_pose else None
        intr_bar = torch.empty(B, 3, 3, device=dev) if need_intr else None
        ws = torch.empty(B * 18, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.sc_pixel_rays_backward(_lib.ptr(p), _lib.ptr(k), _lib.ptr(idx if has_idx else None), B, R, W,
                                                _lib.ptr(g_center), _lib.ptr(g_dirs), _lib.ptr(g_fac), _lib.ptr(ws),
                                                _lib.ptr(pose_bar), _lib.ptr(intr_bar), _lib.stream_of(p)), "sc_pixel_rays_backward")
        rn.TIMERS.count(2)
        return pose_bar, intr_bar, None, None, None


class _EikonalPoints(torch.autograd.Function):
    """cat(uniform points, cam_loc + z_eik * ray_dirs) [B*2R,3] (model/renderer.py:154-170), one launch each way."""

    @staticmethod
    def forward(ctx, cam_loc, ray_dirs, scale_dist, t_vals, jitter, eik_idx, uniform_pts, cam_dist, half_range):
        from . import _lib, _render_native as rn
        L = _lib.lib()
        _lib.require_cuda(cam_loc, ray_dirs, scale_dist, t_vals, jitter, eik_idx, uniform_pts)
        B, R = ray_dirs.shape[0], ray_dirs.shape[1]
        S = t_vals.shape[0]
        dev = ray_dirs.device
        c = lambda t: rn._f32c(t.detach()) if t is not None else None
        loc, dirs, sd, tv, jit, uni = c(cam_loc), c(ray_dirs), c(scale_dist), c(t_vals), c(jitter), c(uniform_pts)
        idx = eik_idx.contiguous()
        pts = torch.empty(B, 2 * R, 3, device=dev); z = torch.empty(B, R, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.sc_eikonal_points_forward(_lib.ptr(loc), _lib.ptr(dirs), _lib.ptr(sd), _lib.ptr(tv), _lib.ptr(jit), _lib.ptr(idx),
                                                   _lib.ptr(uni), B, R, S, float(cam_dist), float(half_range), _lib.ptr(pts), _lib.ptr(z),
                                                   _lib.stream_of(dirs)), "sc_eikonal_points_forward")
        rn.TIMERS.count()
        ctx.save_for_backward(dirs, z)
        ctx.cam_dist = float(cam_dist)
        return pts.reshape(-1, 3)

    @staticmethod
    def backward(ctx, g_pts):
        from . import _lib, _render_native as rn
        L = _lib.lib()
        dirs, z = ctx.saved_tensors
        B, R = dirs.shape[0], dirs.shape[1]
        dev = dirs.device
        g = rn._f32c(g_pts)
        dirs_bar = torch.empty(B, R, 3, device=dev); acc = torch.empty(B, 4, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.sc_eikonal_points_backward(_lib.ptr(dirs), _lib.ptr(z), _lib.ptr(g), B, R, ctx.cam_dist, _lib.ptr(dirs_bar),
                                                    _lib.ptr(acc), _lib.stream_of(dirs)), "sc_eikonal_points_backward")
        rn.TIMERS.count()
        return acc[:, :3], dirs_bar, acc[:, 3], None, None, None, None, None, None


def eikonal_points(cam_loc, ray_dirs, scale_dist, t_vals, jitter, eik_idx, uniform_pts, cam_dist, half_range=0.7):
    """[B*2R,3] eikonal sample points of one render (fused CUDA path)."""
    return _EikonalPoints.apply(cam_loc, ray_dirs, scale_dist, t_vals, jitter, eik_idx, uniform_pts, cam_dist, half_range)


def pixel_rays_torch(pose, intr, H, W, ray_idx=None):
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
