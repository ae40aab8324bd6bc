"""Losses that consume the renderer's outputs, with the reference's method names and argument meaning
(model/loss.py:9-97; combination as in model/graph.py:220-265 and model/runner.py:294-305).

Same values as the reference, but written without device->host synchronisation: the trimmed ("robust") normal
loss ranks the masked angular errors instead of boolean-indexing and slicing, so a training step never stalls the
stream. The camera losses of the reference only involve its CNN view estimator and stay there.
"""
import torch
import torch.nn as nn


class Loss(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.normal_l1 = float(opt.reg.normal_l1)
        self.mask_mse = float(opt.reg.mask_mse)

    @staticmethod
    def aggregate_loss(loss, weight=None):
        return (loss * weight).mean() if weight is not None else loss.mean()

    def L1_loss(self, pred, label=0, weight=None):
        return self.aggregate_loss((pred - label).abs(), weight)

    def MSE_loss(self, pred, label=0, weight=None, tolerance=0.):
        if tolerance > 1.e-5:
            raise NotImplementedError("trimmed MSE is not used by the reference configuration")
        return self.aggregate_loss((pred - label) ** 2, weight)

    def iou_loss(self, inputs, targets, weight=None, tolerance=0.):
        if tolerance > 1.e-5:
            raise NotImplementedError("trimmed IoU is not used by the reference configuration")
        B = inputs.shape[0]
        a, b = inputs.reshape(B, -1), targets.reshape(B, -1)
        loss = 1 - (a * b).sum(1) / (a + b - a * b + 1.e-8).sum(1)
        if weight is not None:
            loss = loss * weight.reshape(B)
        return loss.mean()

    def mask_loss(self, inputs, targets, weight=None, tolerance=0.):
        out = self.iou_loss(inputs, targets, weight=weight, tolerance=tolerance)
        if self.mask_mse != 0.:
            out = out + self.mask_mse * self.MSE_loss(inputs, targets, weight=weight, tolerance=tolerance)
        return out

    def normal_loss(self, normal_pred, normal_gt, mask, weight=None, tolerance=0.):
        """Mean of (normal_l1 * L1 + angular) over the int(n * (1 - tolerance)) masked pixels with the smallest
        angular error (n = number of masked pixels), as model/loss.py:52-67."""
        valid = mask.reshape(normal_pred.shape[:2])
        ang = 1 - (normal_pred * normal_gt).sum(-1)
        per_px = self.normal_l1 * (normal_pred - normal_gt).abs().sum(-1) + ang
        if weight is not None:
            per_px = per_px * weight.expand_as(normal_pred)[..., 0]
        key = torch.where(valid, ang.detach(), torch.full_like(ang, float("inf"))).reshape(-1)
        order = torch.sort(key, dim=0, descending=False, stable=True)[1]
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.numel(), device=order.device)
        n_keep = torch.floor(valid.sum().double() * (1 - tolerance)).long()
        keep = (rank < n_keep).reshape(valid.shape) & valid
        return (per_px * keep).sum() / n_keep


class _FusedRenderLosses(torch.autograd.Function):
    """(rgb, mask, normal, eik | None, rgb_t, mask_t, normal_t) -> losses[4] = render MSE, soft-IoU mask loss, trimmed normal
    loss, eikonal MSE, with their input gradients, in two kernel launches + one sort (sc_render_losses_pass1 / _pass2)
    instead of ~130 torch launches per render."""

    @staticmethod
    def forward(ctx, rgb, mask, normal, eik, rgb_t, mask_t, normal_t, normal_l1, normal_tol):
        from . import _lib, _render_native as rn
        L = _lib.lib()
        _lib.require_cuda(rgb, mask, normal, eik, rgb_t, mask_t, normal_t)
        dev = rgb.device
        B, R = rgb.shape[0], rgb.shape[1]
        c = lambda t: rn._f32c(t.detach()) if t is not None else None
        rgb_c, mask_c, normal_c, eik_c, rgb_tc, mask_tc, normal_tc = (c(t) for t in (rgb, mask, normal, eik, rgb_t, mask_t, normal_t))
        n_eik = eik_c.numel() if eik_c is not None else 0
        ws = torch.empty(L.sc_render_losses_workspace_floats(B), device=dev)
        key = torch.empty(B * R, device=dev); per_px = torch.empty(B * R, device=dev)
        rgb_u = torch.empty(B, R, 3, device=dev); mask_u = torch.empty_like(mask_c)
        normal_u = torch.empty(B, R, 3, device=dev); normal_tu = torch.empty(B, R, 3, device=dev)
        eik_u = torch.empty_like(eik_c) if eik_c is not None else None
        losses = torch.empty(4, device=dev)
        p = _lib.ptr
        with torch.cuda.device(dev):
            _lib.check(L.sc_render_losses_pass1(p(rgb_c), p(rgb_tc), p(mask_c), p(mask_tc), p(normal_c), p(normal_tc), p(eik_c), n_eik,
                                                B, R, float(normal_l1), p(ws), p(key), p(per_px), p(rgb_u), p(normal_u), p(normal_tu),
                                                p(eik_u), _lib.stream_of(rgb_c)), "sc_render_losses_pass1")
            order = torch.sort(key, dim=0, descending=False, stable=True)[1]
            _lib.check(L.sc_render_losses_pass2(p(mask_c), p(mask_tc), p(normal_c), p(normal_tc), p(order), p(per_px), n_eik, B, R,
                                                float(normal_l1), float(normal_tol), 0.0, p(ws), p(mask_u), p(normal_u), p(normal_tu),
                                                p(losses), _lib.stream_of(rgb_c)), "sc_render_losses_pass2")
        rn.TIMERS.count(3)
        ctx.save_for_backward(rgb_u, mask_u, normal_u, normal_tu, eik_u if eik_u is not None else losses)
        ctx.has_eik = eik_u is not None
        ctx.shapes = (rgb.shape, mask.shape, normal.shape, eik.shape if eik is not None else None, normal_t.shape)
        ctx.set_materialize_grads(False)
        return tuple(losses.unbind(0))                 # four 0-dim tensors: each loss gets its own upstream gradient

    @staticmethod
    def backward(ctx, g_render, g_mask_l, g_normal_l, g_eik_l):
        rgb_u, mask_u, normal_u, normal_tu, eik_u = ctx.saved_tensors
        s_rgb, s_mask, s_normal, s_eik, s_nt = ctx.shapes
        need = ctx.needs_input_grad
        g_rgb = (rgb_u * g_render).view(s_rgb) if (need[0] and g_render is not None) else None
        g_mask = (mask_u * g_mask_l).view(s_mask) if (need[1] and g_mask_l is not None) else None
        g_normal = (normal_u * g_normal_l).view(s_normal) if (need[2] and g_normal_l is not None) else None
        g_eik = (eik_u * g_eik_l).view(s_eik) if (ctx.has_eik and need[3] and g_eik_l is not None) else None
        g_nt = (normal_tu * g_normal_l).view(s_nt) if (need[6] and g_normal_l is not None) else None
        return g_rgb, g_mask, g_normal, g_eik, None, None, g_nt, None, None


def fused_render_losses(loss_fns, rgb, mask, normal, eik, rgb_t, mask_t, normal_t, normal_tol):
    """-> dict(render, mask, normal[, eikonal]) with the reference's loss definitions (model/loss.py), fused on CUDA tensors."""
    if not rgb.is_cuda or loss_fns.mask_mse != 0.:
        raise NotImplementedError("fused losses need CUDA tensors and reg.mask_mse == 0")
    out = _FusedRenderLosses.apply(rgb, mask, normal, eik, rgb_t, mask_t, normal_t, loss_fns.normal_l1, float(normal_tol))
    d = dict(render=out[0], mask=out[1], normal=out[2])
    if eik is not None:
        d["eikonal"] = out[3]
    return d


def render_losses(loss_fns, opt, out, target, prefix=""):
    """out/target: dicts with rgb, mask, normal (+ grad_eikonal in out). Returns the reference's loss names."""
    L = {}
    L[prefix + "render" if not prefix else "nearest_img"] = loss_fns.MSE_loss(out["rgb"], target["rgb"])
    L[prefix + "mask" if not prefix else "nearest_mask"] = loss_fns.mask_loss(out["mask"], target["mask"])
    valid = (target["mask"] > 0.5) & (out["mask"] > 0.5)
    L[prefix + "normal" if not prefix else "nearest_normal"] = loss_fns.normal_loss(
        out["normal"], target["normal"], valid, tolerance=opt.reg.normal_tol)
    if not prefix and out.get("grad_eikonal") is not None:
        B = out["rgb"].shape[0]
        L["eikonal"] = loss_fns.MSE_loss(out["grad_eikonal"].view(B, -1), 1)
    return L


def summarize_loss(opt, loss):
    """Weighted sum (model/runner.py:294-305) without the per-key isinf/isnan host syncs; NaN/Inf surfaces in `all`."""
    total = 0.
    for key, value in loss.items():
        w = opt.loss_weight.get(key) if hasattr(opt.loss_weight, "get") else getattr(opt.loss_weight, key, None)
        if w is not None:
            total = total + float(w) * value
    return total
