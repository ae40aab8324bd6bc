"""torch.autograd glue between the reference-shaped modules (renderer.py / implicit.py) and the C ABI
(sc_render_forward / sc_render_backward / sc_render_grad_finalize). All arithmetic happens in the CUDA library."""
import ctypes

import torch

from . import _lib
from . import _render_native as rn

_N_SDF, _N_RGB = 6, 4

# Which generation of the render kernels runs: "tc2" / "tc" = tcgen05 tensor cores on hi/lo bf16 operand pairs (3 MMAs per
# product, ~1e-5 relative; "tc2" = two independent 64-point tile chains per CTA, "tc" = one 128-point tile per CTA),
# "bf16" = the "tc" kernels with ONE MMA per product on plain bf16 operands (ScRenderArgs.precision = 1; BASELINE.json
# configs[2]'s arithmetic, ~1e-2 relative), "fp32" = FP32 FFMA (the bit-for-bit-closest path). All are CUDA kernels of this library.
import os as _os
PRECISION = {"forward": _os.environ.get("SC_RENDER_FORWARD", "tc"), "backward": _os.environ.get("SC_RENDER_BACKWARD", "tc")}


STEP_PRECISIONS = ("tc", "bf16")   # render arithmetic modes bench.py's configs[2] record steps through


SAVE_ACTIVATIONS = _os.environ.get("SC_RENDER_SAVE_ACTIVATIONS", "1") != "0"
# A render too large to keep its activations (3.75 KB per sample point; e.g. 128 x 128 rays x 64 images = 252 GB): the backward walks
# the batch in chunks of images — forward again WITH the saved buffer for the chunk, then the saved-activation backward on it —
# instead of the kernel that recomputes the forward per tile (1.3 + 2.6 ns per sample point against 5.4).
CHUNKED_BACKWARD = _os.environ.get("SC_RENDER_CHUNKED_BACKWARD", "1") != "0"
# > 0: never keep a whole render's activations; the backward always walks chunks of this many images (bounds the memory of a
# training render to chunk x 3.75 KB x sample points per image; also how the tests reach the chunked path at small sizes)
CHUNK_IMAGES = int(_os.environ.get("SC_RENDER_CHUNK_IMAGES", "0"))


def set_precision(forward=None, backward=None):
    for k, v in (("forward", forward), ("backward", backward)):
        if v is not None:
            if v not in ("tc2", "tc", "bf16", "fp32"):
                raise ValueError("precision must be 'tc2', 'tc', 'bf16' or 'fp32'")
            PRECISION[k] = v


def _fwd_tc():
    """False (FP32 FFMA) or the tensor-core generation name ("tc" / "tc2", both truthy); "bf16" runs the "tc" generation."""
    p = PRECISION["forward"]
    return ("tc" if p == "bf16" else p) if p in ("tc", "tc2", "bf16") else False


def _bwd_tc():
    p = PRECISION["backward"]
    return ("tc" if p == "bf16" else p) if p in ("tc", "tc2", "bf16") else False


def _single(direction):
    """ScRenderArgs.precision: 1 = one MMA per product (plain bf16 operands), 0 = three (hi/lo pairs)."""
    return 1 if PRECISION[direction] == "bf16" else 0


def _params_of(sdf_net, rgb_net, device):
    ws = [l.weight for l in sdf_net.linears()]
    bs = [l.bias for l in sdf_net.linears()]
    if rgb_net is not None:
        ws += [l.weight for l in rgb_net.linears()]
        bs += [l.bias for l in rgb_net.linears()]
    return ws, bs


_dummy_rgb = {}


def _dummy_rgb_params(device):
    key = str(device)
    if key not in _dummy_rgb:
        shapes = [(64, 167), (64, 64), (64, 64), (3, 64)]
        _dummy_rgb[key] = ([torch.zeros(s, device=device) for s in shapes],
                           [torch.zeros(s[0], device=device) for s in shapes])
    return _dummy_rgb[key]


def _new_args(**kw):
    a = rn.ScRenderArgs()
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = ctypes.c_void_p(v.data_ptr())
        setattr(a, k, v)
    return a


# Fused gradient accumulation: the finalize kernel adds the parameter gradients straight into existing `.grad` tensors and
# the autograd Function returns None for them — one launch instead of one accumulation kernel per parameter and call
# (~70 launches per training step). OFF by default: it bypasses the parameters' AccumulateGrad nodes, so hooks on them
# (e.g. the reference's DistributedDataParallel) would not fire. step.TrainStep, which owns the flat gradient buffer and its
# all-reduce, switches it on around its forward/backward.
FUSED_GRAD_ACCUMULATION = False


def _finalize(L, partial, n_ctas, cb_bar, z_sdf, z_rgb, blob, B, ws, bs, need_rgb, needs=None):
    """-> (grads for weights[10], biases[10] (None where not produced), z_sdf_bar, z_rgb_bar, beta_eff_bar).
    needs: per-tensor needs_input_grad flags (weights then biases) of the calling Function, or None."""
    dev = blob.device
    n_w = 10 if need_rgb else _N_SDF
    live = list(ws[:n_w]) + list(bs[:n_w])
    fused = FUSED_GRAD_ACCUMULATION and all(
        (p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous() and p.grad.device == dev)
        for p in live)
    if fused:
        gw = [w.grad for w in ws[:n_w]] + [None] * (10 - n_w)
        gb = [b.grad for b in bs[:n_w]] + [None] * (10 - n_w)
    else:
        gw = [torch.empty_like(w) for w in ws[:n_w]] + [None] * (10 - n_w)
        gb = [torch.empty_like(b) for b in bs[:n_w]] + [None] * (10 - n_w)
    warr = (ctypes.c_void_p * 10)(*[(g.data_ptr() if g is not None else None) for g in gw])
    barr = (ctypes.c_void_p * 10)(*[(g.data_ptr() if g is not None else None) for g in gb])
    z_sdf_bar = torch.empty(B, 64, device=dev)
    z_rgb_bar = torch.empty(B, 64, device=dev) if need_rgb else None
    beta_bar = torch.empty(1, device=dev)
    fn = L.sc_render_grad_finalize_accumulate if fused else L.sc_render_grad_finalize
    with torch.cuda.device(dev):
        _lib.check(fn(rn._p(partial), n_ctas, rn._p(cb_bar), rn._p(z_sdf), rn._p(z_rgb), rn._p(blob), B,
                      warr, barr, rn._p(z_sdf_bar), rn._p(z_rgb_bar), rn._p(beta_bar),
                      _lib.stream_of(blob)), "sc_render_grad_finalize")
    rn.TIMERS.count()
    if fused:                                     # already in p.grad: nothing for autograd to accumulate
        gw, gb = [None] * 10, [None] * 10
    return gw, gb, z_sdf_bar, z_rgb_bar, beta_bar


class _RenderFn(torch.autograd.Function):
    """Renderer.forward minus ray generation and the eikonal branch (model/renderer.py:80-152)."""

    @staticmethod
    def forward(ctx, cfg, beta_param, cam_loc, ray_dirs, depth_fac, scale_dist, z_sdf, z_rgb, t_vals, jitter, *params):
        L = _lib.lib()
        _lib.require_cuda(cam_loc, ray_dirs, depth_fac, scale_dist, z_sdf, z_rgb, beta_param, *params)
        dev = ray_dirs.device
        B, R = ray_dirs.shape[0], ray_dirs.shape[1]
        ws, bs = list(params[:10]), list(params[10:])
        blob = rn.packed_blob(ws, bs)
        cb = rn.latent_bias(blob, z_sdf, z_rgb, B)
        f = lambda t: rn._f32c(t.detach())
        cam_loc, ray_dirs, depth_fac, scale_dist = f(cam_loc), f(ray_dirs), f(depth_fac), f(scale_dist)
        jit = f(jitter) if jitter is not None else None
        t_vals = f(t_vals)
        rgb = torch.empty(B, R, 3, device=dev); normal = torch.empty(B, R, 3, device=dev)
        mask = torch.empty(B, R, 1, device=dev); mask_hard = torch.empty(B, R, 1, device=dev)
        depth = torch.empty(B, R, 1, device=dev)
        tc = _fwd_tc()
        scratch = rn.scratch(dev, backward=False, tc=tc)
        beta_c = f(beta_param).reshape(1)
        kblob = rn.packed_tc_blob(ws, bs, blob) if tc else blob
        args = _new_args(mode=0, batch=B, n_per_image=R, n_samples=cfg["n_samples"], beta_min=cfg["beta_min"],
                         cam_dist=cfg["cam_dist"], half_range=cfg["half_range"], bg_color=cfg["bg_color"],
                         normal_pow=cfg["normal_pow"], blob=kblob, cb=cb, beta_param=beta_c, cam_loc=cam_loc,
                         ray_dirs=ray_dirs, depth_fac=depth_fac, scale_dist=scale_dist, t_vals=t_vals,
                         rgb=rgb, mask=mask, mask_hard=mask_hard, depth=depth, normal=normal, scratch=scratch,
                         precision=_single("forward"))
        if jit is not None:
            args.jitter = ctypes.c_void_p(jit.data_ptr())
        # generation-1 tensor-core kernels, some input needs a gradient: the forward saves its per-point activations
        # (3.75 KB per sample point) and the backward reads them back instead of recomputing 19 of its 45 GEMM phases
        saved = None
        if tc == "tc" and _bwd_tc() == "tc" and SAVE_ACTIVATIONS and any(ctx.needs_input_grad) and not (CHUNK_IMAGES > 0 and CHUNKED_BACKWARD):
            saved = rn.saved_buffer(dev, B, R, cfg["n_samples"])
            if saved is not None:
                args.saved = ctypes.c_void_p(saved.data_ptr())
        rn.launch_forward(args, dev, tc=tc)
        ctx.cfg = cfg
        ctx.has_jitter = jit is not None
        ctx.saved_acts = saved
        ctx.save_for_backward(blob, cb, beta_c, cam_loc, ray_dirs, depth_fac, scale_dist, t_vals,
                              jit if jit is not None else t_vals, f(z_sdf), f(z_rgb), *params)
        ctx.mark_non_differentiable(mask_hard)
        ctx.set_materialize_grads(False)
        return rgb, mask, mask_hard, depth, normal

    @staticmethod
    def backward(ctx, rgb_bar, mask_bar, _mh_bar, depth_bar, normal_bar):
        L = _lib.lib()
        saved = ctx.saved_tensors
        blob, cb, beta_c, cam_loc, ray_dirs, depth_fac, scale_dist, t_vals, jit, z_sdf, z_rgb = saved[:11]
        params = saved[11:]
        ws, bs = list(params[:10]), list(params[10:])
        cfg = ctx.cfg
        dev = ray_dirs.device
        B, R = ray_dirs.shape[0], ray_dirs.shape[1]
        g = lambda t: rn._f32c(t) if t is not None else None
        rgb_bar, mask_bar, depth_bar, normal_bar = g(rgb_bar), g(mask_bar), g(depth_bar), g(normal_bar)
        n_ctas = L.sc_render_num_ctas()
        partial = torch.empty(n_ctas, L.sc_render_grad_floats(), device=dev)
        zeros = torch.zeros(B * (7 * 64 + R * 4 + 4), device=dev)          # one fill for every atomically accumulated output
        o = 0
        cb_bar = zeros[o:o + B * 448].view(B, 7, 64); o += B * 448
        dirs_bar = zeros[o:o + B * R * 3].view(B, R, 3); o += B * R * 3
        fac_bar = zeros[o:o + B * R].view(B, R); o += B * R
        loc_bar = zeros[o:o + B * 3].view(B, 3); o += B * 3
        sd_bar = zeros[o:o + B]
        tc = _bwd_tc()
        scratch = rn.scratch(dev, backward=True, tc=tc)
        kblob = rn.packed_tc_blob(ws, bs, blob) if tc else blob
        args = _new_args(mode=0, batch=B, n_per_image=R, n_samples=cfg["n_samples"], beta_min=cfg["beta_min"],
                         cam_dist=cfg["cam_dist"], half_range=cfg["half_range"], bg_color=cfg["bg_color"],
                         normal_pow=cfg["normal_pow"], blob=kblob, cb=cb, beta_param=beta_c, cam_loc=cam_loc,
                         ray_dirs=ray_dirs, depth_fac=depth_fac, scale_dist=scale_dist, t_vals=t_vals,
                         grad_partial=partial, cb_bar=cb_bar, ray_dirs_bar=dirs_bar, depth_fac_bar=fac_bar,
                         cam_loc_bar=loc_bar, scale_dist_bar=sd_bar, scratch=scratch, precision=_single("backward"))
        if ctx.has_jitter:
            args.jitter = ctypes.c_void_p(jit.data_ptr())
        for name, t in (("rgb_bar", rgb_bar), ("mask_bar", mask_bar), ("depth_bar", depth_bar), ("normal_bar", normal_bar)):
            if t is not None:
                setattr(args, name, ctypes.c_void_p(t.data_ptr()))
        S = cfg["n_samples"]
        n_img = 0
        if (ctx.saved_acts is None and tc == "tc" and _fwd_tc() == "tc" and SAVE_ACTIVATIONS and CHUNKED_BACKWARD and S <= 64
                and _single("forward") == _single("backward")):
            n_img = min(CHUNK_IMAGES, B) if CHUNK_IMAGES > 0 else rn.saved_chunk_images(dev, B, R, S)
        buf = rn.saved_buffer(dev, n_img, R, S) if (n_img >= 1 and (B > 1 or CHUNK_IMAGES > 0)) else None
        if buf is not None:                        # (None: no memory for even one chunk -> the per-tile recompute kernel below)
            return _RenderFn._backward_chunked(ctx, L, args, n_img, buf, (rgb_bar, mask_bar, depth_bar, normal_bar),
                                               (cb_bar, dirs_bar, fac_bar, loc_bar, sd_bar), partial, n_ctas)
        if ctx.saved_acts is not None and tc == "tc":
            args.saved = ctypes.c_void_p(ctx.saved_acts.data_ptr())
        rn.launch_backward(args, dev, tc=tc)
        ctx.saved_acts = None
        gw, gb, z_sdf_bar, z_rgb_bar, beta_bar = _finalize(L, partial, n_ctas, cb_bar, z_sdf, z_rgb, blob, B, ws, bs, True)
        beta_param_bar = (beta_bar * torch.sign(beta_c)).reshape(())
        return (None, beta_param_bar, loc_bar, dirs_bar, fac_bar, sd_bar, z_sdf_bar, z_rgb_bar, None, None, *gw, *gb)

    @staticmethod
    def _backward_chunked(ctx, L, args, n_img, buf, ups, outs, partial, n_ctas):
        """The backward of a render whose activations did not fit: per chunk of n_img images, the forward kernel again with the
        saved-activation buffer (outputs discarded: same inputs, same jitter => same planes), the saved-activation backward kernel,
        and one finalize that adds the chunk's parameter gradients to the running sums."""
        saved = ctx.saved_tensors
        blob, cb, beta_c, cam_loc, ray_dirs, depth_fac, scale_dist, t_vals, jit, z_sdf, z_rgb = saved[:11]
        params = saved[11:]
        ws, bs = list(params[:10]), list(params[10:])
        cfg = ctx.cfg
        dev = ray_dirs.device
        B, R, S = ray_dirs.shape[0], ray_dirs.shape[1], cfg["n_samples"]
        cb_bar, dirs_bar, fac_bar, loc_bar, sd_bar = outs
        dummy = [torch.empty(n_img, R, c, device=dev) for c in (3, 1, 1, 1, 3)]         # rgb mask mask_hard depth normal of a chunk
        fscratch = rn.scratch(dev, backward=False, tc="tc")
        kblob = rn.packed_tc_blob(ws, bs, blob)
        addr = lambda t, off: ctypes.c_void_p(t.data_ptr() + 4 * int(off))
        gw_sum = gb_sum = None
        z_sdf_bar = torch.empty(B, 64, device=dev)
        z_rgb_bar = torch.empty(B, 64, device=dev)
        beta_sum = torch.zeros(1, device=dev)
        for b0 in range(0, B, n_img):
            nb = min(n_img, B - b0)
            fa = _new_args(mode=0, batch=nb, n_per_image=R, n_samples=S, beta_min=cfg["beta_min"], cam_dist=cfg["cam_dist"],
                           half_range=cfg["half_range"], bg_color=cfg["bg_color"], normal_pow=cfg["normal_pow"], blob=kblob,
                           beta_param=beta_c, t_vals=t_vals, rgb=dummy[0], mask=dummy[1], mask_hard=dummy[2], depth=dummy[3],
                           normal=dummy[4], scratch=fscratch, saved=buf, precision=_single("forward"))
            fa.cb = addr(cb, b0 * 256); fa.cam_loc = addr(cam_loc, b0 * 3); fa.ray_dirs = addr(ray_dirs, b0 * R * 3)
            fa.depth_fac = addr(depth_fac, b0 * R); fa.scale_dist = addr(scale_dist, b0)
            if ctx.has_jitter:
                fa.jitter = addr(jit, b0 * R * S)
            rn.launch_forward(fa, dev, tc="tc", span="render_fwd_for_backward")
            ba = _new_args(mode=0, batch=nb, n_per_image=R, n_samples=S, beta_min=cfg["beta_min"], cam_dist=cfg["cam_dist"],
                           half_range=cfg["half_range"], bg_color=cfg["bg_color"], normal_pow=cfg["normal_pow"], blob=kblob,
                           beta_param=beta_c, t_vals=t_vals, grad_partial=partial, scratch=args.scratch, saved=buf,
                           precision=_single("backward"))
            ba.cb = fa.cb; ba.cam_loc = fa.cam_loc; ba.ray_dirs = fa.ray_dirs; ba.depth_fac = fa.depth_fac; ba.scale_dist = fa.scale_dist
            if ctx.has_jitter:
                ba.jitter = fa.jitter
            for name, t, w in (("rgb_bar", ups[0], 3), ("mask_bar", ups[1], 1), ("depth_bar", ups[2], 1), ("normal_bar", ups[3], 3)):
                if t is not None:
                    setattr(ba, name, addr(t, b0 * R * w))
            ba.cb_bar = addr(cb_bar, b0 * 448); ba.ray_dirs_bar = addr(dirs_bar, b0 * R * 3); ba.depth_fac_bar = addr(fac_bar, b0 * R)
            ba.cam_loc_bar = addr(loc_bar, b0 * 3); ba.scale_dist_bar = addr(sd_bar, b0)
            rn.launch_backward(ba, dev, tc="tc")
            gw, gb, zs, zr, beta_bar = _finalize(L, partial, n_ctas, cb_bar[b0:b0 + nb], z_sdf[b0:b0 + nb], z_rgb[b0:b0 + nb], blob, nb,
                                                 ws, bs, True)
            z_sdf_bar[b0:b0 + nb] = zs
            z_rgb_bar[b0:b0 + nb] = zr
            beta_sum += beta_bar
            if gw[0] is not None:                                   # not fused into .grad: sum the chunks here
                if gw_sum is None:
                    gw_sum, gb_sum = list(gw), list(gb)
                else:
                    for i in range(10):
                        if gw[i] is not None:
                            gw_sum[i] += gw[i]; gb_sum[i] += gb[i]
        if gw_sum is None:
            gw_sum, gb_sum = [None] * 10, [None] * 10
        beta_param_bar = (beta_sum * torch.sign(beta_c)).reshape(())
        return (None, beta_param_bar, loc_bar, dirs_bar, fac_bar, sd_bar, z_sdf_bar, z_rgb_bar, None, None, *gw_sum, *gb_sum)


class _SDFQueryFn(torch.autograd.Function):
    """SDFNetwork.get_conditional_output (model/implicit.py:163-189): sdf, features and d sdf / d x."""

    @staticmethod
    def forward(ctx, B, want_grad, detach_latent, points, z_sdf, *params):
        L = _lib.lib()
        _lib.require_cuda(points, z_sdf, *params)
        dev = points.device
        P = points.shape[0]
        if P % B != 0:
            raise ValueError("points_flat must hold batch_size * N points (batch-major)")
        N = P // B
        dw, db = _dummy_rgb_params(dev)
        ws = list(params[:6]) + dw
        bs = list(params[6:]) + db
        blob = rn.packed_blob(ws, bs)
        cb = rn.latent_bias(blob, z_sdf, None, B)
        pts = rn._f32c(points.detach())
        sdf = torch.empty(P, 1, device=dev); feat = torch.empty(P, 64, device=dev)
        grad = torch.empty(P, 3, device=dev) if want_grad else None
        tc = _fwd_tc()
        scratch = rn.scratch(dev, backward=False, tc=tc)
        kblob = rn.packed_tc_blob(ws, bs, blob) if tc else blob
        args = _new_args(mode=1, batch=B, n_per_image=N, n_samples=1, want_grad=int(want_grad), want_feat=1,
                         beta_min=1e-4, blob=kblob, cb=cb, points=pts, sdf=sdf, feat=feat, scratch=scratch, precision=_single("forward"))
        if grad is not None:
            args.grad = ctypes.c_void_p(grad.data_ptr())
        rn.launch_forward(args, dev, tc=tc)
        ctx.meta = (B, N, want_grad, detach_latent)
        ctx.save_for_backward(blob, cb, pts, rn._f32c(z_sdf.detach()), *params)
        ctx.set_materialize_grads(False)
        if grad is None:
            grad = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(grad)
        return sdf, feat, grad

    @staticmethod
    def backward(ctx, sdf_bar, feat_bar, grad_bar):
        if feat_bar is not None:
            raise NotImplementedError("gradients through the SDF feature output of get_conditional_output are only "
                                      "supported inside Renderer.forward (the reference never needs them elsewhere)")
        L = _lib.lib()
        blob, cb, pts, z_sdf = ctx.saved_tensors[:4]
        params = ctx.saved_tensors[4:]
        B, N, want_grad, detach_latent = ctx.meta
        dev = pts.device
        dw, db = _dummy_rgb_params(dev)
        ws, bs = list(params[:6]) + dw, list(params[6:]) + db
        n_ctas = L.sc_render_num_ctas()
        partial = torch.empty(n_ctas, L.sc_render_grad_floats(), device=dev)
        cb_bar = torch.zeros(B, 7, 64, device=dev)
        pts_bar = torch.zeros(B * N, 3, device=dev)
        tc = _bwd_tc()
        scratch = rn.scratch(dev, backward=True, tc=tc)
        kblob = rn.packed_tc_blob(ws, bs, blob) if tc else blob
        args = _new_args(mode=1, batch=B, n_per_image=N, n_samples=1, want_grad=int(want_grad and grad_bar is not None),
                         want_feat=0, detach_latent=int(detach_latent), beta_min=1e-4, blob=kblob, cb=cb, points=pts,
                         grad_partial=partial, cb_bar=cb_bar, points_bar=pts_bar, scratch=scratch, precision=_single("backward"))
        if sdf_bar is not None:
            keep1 = rn._f32c(sdf_bar)
            args.sdf_bar = ctypes.c_void_p(keep1.data_ptr())
        if grad_bar is not None and want_grad:
            keep2 = rn._f32c(grad_bar)
            args.grad_bar = ctypes.c_void_p(keep2.data_ptr())
        rn.launch_backward(args, dev, tc=tc)
        gw, gb, z_sdf_bar, _, _ = _finalize(L, partial, n_ctas, cb_bar, z_sdf, None, blob, B, ws, bs, False)
        return (None, None, None, pts_bar, (None if detach_latent else z_sdf_bar), *gw[:6], *gb[:6])


def sdf_query(sdf_net, opt, batch_size, points_flat, proj_latent, want_grad, detach_latent):
    ws = [l.weight for l in sdf_net.linears()]
    bs = [l.bias for l in sdf_net.linears()]
    sdf, feat, grad = _SDFQueryFn.apply(int(batch_size), bool(want_grad), bool(detach_latent), points_flat, proj_latent,
                                        *ws, *bs)
    return sdf, feat, (grad if want_grad else None)


def render_rays(cfg, beta_param, cam_loc, ray_dirs, depth_fac, scale_dist, z_sdf, z_rgb, t_vals, jitter, sdf_net, rgb_net):
    ws = [l.weight for l in sdf_net.linears()] + [l.weight for l in rgb_net.linears()]
    bs = [l.bias for l in sdf_net.linears()] + [l.bias for l in rgb_net.linears()]
    return _RenderFn.apply(cfg, beta_param, cam_loc, ray_dirs, depth_fac, scale_dist, z_sdf, z_rgb, t_vals, jitter, *ws, *bs)
