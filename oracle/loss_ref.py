"""ORACLE (test infrastructure): the losses that consume renderer outputs.

Restates model/loss.py:19-97 (MSE_loss, normal_loss, iou_loss, mask_loss) and the part of
model/graph.py:220-265 + model/runner.py:294-305 that combines them. The camera losses
(cam_margin / cam_uniform / cam_sym) only touch the CNN view estimator and are out of scope.
"""
import torch


def mse(pred, label=0.0):
    return ((pred - label) ** 2).mean()


def soft_iou(pred, target):
    B = pred.shape[0]
    a, b = pred.reshape(B, -1), target.reshape(B, -1)
    inter = (a * b).sum(1)
    union = (a + b - a * b + 1.0e-8).sum(1)
    return (1 - inter / union).mean()


def mask_loss(pred, target, mask_mse=0.0):
    return soft_iou(pred, target) + mask_mse * mse(pred, target)


def trimmed_normal(n_pred, n_gt, valid, tol=0.2, l1_weight=5.0):
    """valid [B,R] bool. Keep the int(n*(1-tol)) pixels with the smallest angular error."""
    p, g = n_pred[valid], n_gt[valid]
    ang = 1 - (p * g).sum(-1)
    per_px = l1_weight * (p - g).abs().sum(-1) + ang
    keep = torch.sort(ang, dim=0)[1][: int(per_px.shape[0] * (1 - tol))]
    return per_px[keep].mean()


def render_losses(out, rgb_gt, mask_gt, normal_gt, batch_size, tol=0.2, l1_weight=5.0, mask_mse=0.0):
    """out = oracle.render_ref.render(...) dict (or any dict with rgb/mask/normal/grad_eik)."""
    L = {}
    L["render"] = mse(out["rgb"], rgb_gt)
    L["mask"] = mask_loss(out["mask"], mask_gt, mask_mse)
    valid = ((mask_gt > 0.5) & (out["mask"] > 0.5)).squeeze(-1)
    L["normal"] = trimmed_normal(out["normal"], normal_gt, valid, tol, l1_weight)
    if out.get("grad_eik") is not None:
        L["eikonal"] = mse(out["grad_eik"].view(batch_size, -1), 1.0)
    return L


DEFAULT_WEIGHTS = dict(render=1.0, mask=0.5, normal=0.01, eikonal=0.03,
                       nearest_img=1.0, nearest_mask=0.5, nearest_normal=0.01)


def weighted_total(L, weights=DEFAULT_WEIGHTS):
    total = 0.0
    for k, v in L.items():
        total = total + float(weights[k]) * v
    return total
