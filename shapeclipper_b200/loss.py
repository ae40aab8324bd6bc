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
