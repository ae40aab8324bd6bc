"""The render half of the reference's Graph (model/graph.py:68-265): 1 + n_views renders per step, the IoU-weighted
CLIP-neighbour selection and the render-consuming losses, with the same `forward(opt, var, training, get_loss)`
calling convention and the same `var` / `loss` key names.

What is NOT here (out of the hot path): the torchvision ResNet encoder and the ResNet view
estimator. Their outputs enter through `var` — `proj_latent_sdf`, `proj_latent_rgb`, `pose`, `intr`, `scale_dist`
for the query image and `proj_latent_rgb_NN [B,64,K]`, `pose_NN [B,3,4,K]`, `intr_NN [B,3,3,K]`,
`scale_dist_NN [B,K]` for its K CLIP neighbours — and gradients flow back to them. With the reference tree present,
`shapeclipper_b200.shim.install()` instead drops the kernels under the reference's own Graph unchanged.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import loss as loss_mod
from .implicit import RGBNetwork, SDFNetwork
from .renderer import Renderer


class HotPathGraph(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.sdf_network = SDFNetwork(opt)
        self.rgb_network = RGBNetwork(opt)
        self.renderer = Renderer(opt, self.sdf_network, self.rgb_network)
        self.loss_fns = loss_mod.Loss(opt)

    # ------------------------------------------------------------------------------------------------------
    def forward(self, opt, var, training=False, get_loss=True, visualize=False):
        ray_idx = var.ray_idx if (opt.render.rand_sample and training) else None
        # canonicalise the normal map: camera-frame normals -> world frame (utils/camera.py:98-103)
        var.normal_transformed = var.normal_input @ var.pose[..., :3]
        (var.rgb_recon, var.mask_recon, var.mask_hard, var.depth_recon, var.normal_recon, var.grad_eikonal) = \
            self.renderer(opt, var.pose, var.intr, var.scale_dist, var.proj_latent_sdf, var.proj_latent_rgb,
                          ray_idx=ray_idx, training=training)
        lw = opt.loss_weight
        if training and (lw.nearest_img is not None or lw.nearest_mask is not None):      # model/graph.py:104
            self.forward_NN(opt, var, training=training)
        if get_loss:
            return var, self.compute_loss(opt, var, training)
        return var

    # ------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def select_neighbours(self, opt, var):
        """IoU(query mask, neighbour mask) -> (1 - IoU)^sample_temp -> L1-normalise -> draw n_views of K without
        replacement (model/graph.py:119-142). `opt.reg.device_sampling` keeps the draw on the GPU (no host sync);
        otherwise numpy's global generator is used exactly as in the reference."""
        B, K, V = var.mask_input.shape[0], opt.data.k_nearest, opt.reg.n_views
        q = var.mask_input.reshape(B, -1, 1)
        nn_ = var.mask_input_NN.reshape(B, -1, K)
        iou = (nn_ * q).sum(1) / (nn_ + q - nn_ * q + 1.e-8).sum(1)
        probs = F.normalize((1 - iou) ** opt.reg.sample_temp, dim=-1, p=1)
        if getattr(opt.reg, "device_sampling", False):
            # V draws without replacement with probabilities `probs`: the V largest of probs / Exp(1) (the exponential-race
            # form torch.multinomial itself uses), without multinomial's validity asserts -> capturable, 3 kernels.
            return (probs / torch.empty_like(probs).exponential_(1.0)).topk(V, dim=-1).indices
        rows = []
        for p in probs.cpu().numpy():
            p = p / np.sum(p)
            rows.append(np.random.choice(K, size=(V,), replace=False, p=p))
        return torch.tensor(np.stack(rows, 0)).long().to(var.mask_input.device)

    def forward_NN(self, opt, var, training=True):
        assert opt.reg.n_views <= opt.data.k_nearest
        idx = self.select_neighbours(opt, var)                       # [B,V]
        var.idx_NN = idx
        B = idx.shape[0]

        def pick(t, v):                                              # t [B,...,K] -> [B,...] at neighbour idx[:, v]
            sel = idx[:, v].reshape(B, *([1] * (t.dim() - 1))).expand(*t.shape[:-1], 1)
            return torch.gather(t, -1, sel).squeeze(-1)
        for v in range(opt.reg.n_views):
            tag = "NN_%d" % v
            inp = dict(rgb_input=pick(var.rgb_input_NN, v), mask_input=pick(var.mask_input_NN, v),
                       normal_input=pick(var.normal_input_NN, v))
            ray_idx = pick(var.ray_idx_NN, v) if (opt.render.rand_sample and training) else None
            pose, intr = pick(var.pose_NN, v), pick(var.intr_NN, v)
            scale_dist, z_rgb = pick(var.scale_dist_NN, v), pick(var.proj_latent_rgb_NN, v)
            var["input_" + tag] = inp
            var["pose_" + tag], var["intr_" + tag], var["scale_dist_" + tag] = pose, intr, scale_dist
            # the QUERY's shape code, the neighbour's appearance code and viewpoint (model/graph.py:207-209)
            rgb, mask, _, depth, normal, _ = self.renderer(opt, pose, intr, scale_dist, var.proj_latent_sdf, z_rgb,
                                                           ray_idx=ray_idx, training=training, eikonal=False)
            var["rgb_recon_" + tag], var["mask_recon_" + tag] = rgb, mask
            var["depth_recon_" + tag], var["normal_recon_" + tag] = depth, normal

    # ------------------------------------------------------------------------------------------------------
    def compute_loss(self, opt, var, training=False):
        lw, fns = opt.loss_weight, self.loss_fns
        nn_w = (lw.nearest_img, lw.nearest_mask, lw.nearest_normal)
        fused = (var.rgb_recon.is_cuda and fns.mask_mse == 0. and lw.render is not None and lw.mask is not None
                 and lw.normal is not None and getattr(opt.reg, "fused_losses", True)
                 and (all(w is not None for w in nn_w) or all(w is None for w in nn_w)))    # the fused kernel computes all three
        if fused:
            return self._compute_loss_fused(opt, var, training)
        L = {}
        if lw.render is not None:
            L["render"] = fns.MSE_loss(var.rgb_recon, var.rgb_input)
        if lw.mask is not None:
            L["mask"] = fns.mask_loss(var.mask_recon, var.mask_input)
        if lw.normal is not None:
            valid = (var.mask_input > 0.5) & (var.mask_recon > 0.5)
            L["normal"] = fns.normal_loss(var.normal_recon, var.normal_transformed, valid, tolerance=opt.reg.normal_tol)
        if lw.eikonal is not None and training:
            L["eikonal"] = fns.MSE_loss(var.grad_eikonal.view(var.rgb_recon.shape[0], -1), 1)
        # each neighbour loss is gated on its OWN weight, as in the reference (model/graph.py:241-263)
        rendered_nn = training and ("rgb_recon_NN_0" in var)
        if rendered_nn and lw.nearest_img is not None:
            L["nearest_img"] = 0
            for v in range(opt.reg.n_views):
                tag = "NN_%d" % v
                L["nearest_img"] = L["nearest_img"] + fns.MSE_loss(var["rgb_recon_" + tag], var["input_" + tag]["rgb_input"])
        if rendered_nn and lw.nearest_mask is not None:
            L["nearest_mask"] = 0
            for v in range(opt.reg.n_views):
                tag = "NN_%d" % v
                L["nearest_mask"] = L["nearest_mask"] + fns.mask_loss(var["mask_recon_" + tag], var["input_" + tag]["mask_input"])
        if rendered_nn and lw.nearest_normal is not None:
            L["nearest_normal"] = 0
            for v in range(opt.reg.n_views):
                tag = "NN_%d" % v
                inp = var["input_" + tag]
                valid = (inp["mask_input"] > 0.5) & (var["mask_recon_" + tag] > 0.5)
                target = inp["normal_input"] @ var["pose_" + tag][..., :3]
                L["nearest_normal"] = L["nearest_normal"] + fns.normal_loss(var["normal_recon_" + tag], target, valid,
                                                                            tolerance=opt.reg.normal_tol)
        L["all"] = loss_mod.summarize_loss(opt, L)
        return L

    def _compute_loss_fused(self, opt, var, training):
        """Same losses through the fused CUDA kernels (loss.fused_render_losses): two launches + one sort per render."""
        lw, fns = opt.loss_weight, self.loss_fns
        eik = var.grad_eikonal if (lw.eikonal is not None and training) else None
        L = loss_mod.fused_render_losses(fns, var.rgb_recon, var.mask_recon, var.normal_recon, eik, var.rgb_input,
                                         var.mask_input, var.normal_transformed, opt.reg.normal_tol)
        if training and lw.nearest_img is not None:
            L["nearest_img"], L["nearest_mask"], L["nearest_normal"] = 0, 0, 0
            for v in range(opt.reg.n_views):
                tag = "NN_%d" % v
                inp = var["input_" + tag]
                target = inp["normal_input"] @ var["pose_" + tag][..., :3]
                N = loss_mod.fused_render_losses(fns, var["rgb_recon_" + tag], var["mask_recon_" + tag], var["normal_recon_" + tag],
                                                 None, inp["rgb_input"], inp["mask_input"], target, opt.reg.normal_tol)
                L["nearest_img"] = L["nearest_img"] + N["render"]
                L["nearest_mask"] = L["nearest_mask"] + N["mask"]
                L["nearest_normal"] = L["nearest_normal"] + N["normal"]
        L["all"] = loss_mod.summarize_loss(opt, L)
        return L
