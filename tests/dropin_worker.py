"""Runs the reference's OWN entry points (model/graph.py Graph, model/runner.py Runner.summarize_loss, utils/eval_3D.py)
either as they are on the CPU (`--mode ref`) or with `shapeclipper_b200.shim.install()` underneath on CUDA (`--mode shim`),
on the same seeded synthetic batch, and dumps losses / outputs / gradients to a .pt file. tests/test_dropin_gpu.py runs it
twice (separate processes, so the two `model.renderer` modules never share an interpreter) and compares the files.

The reference modules come from tests/refharness.py: /root/reference in the build container, the sourceless bytecode
oracle/build_ref.py staged under oracle/_ref/py on the GPU box.
"""
import argparse
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_var(edict, opt, B, seed, device):
    """A batch with every field data/pix3d.py:129-227 produces, maps and ray samples consistent."""
    import torch
    g = torch.Generator().manual_seed(seed)
    H, W = opt.image_size
    K = opt.data.k_nearest
    R = int(opt.render.rand_sample)
    yy, xx = torch.meshgrid(torch.arange(H).float() + 0.5, torch.arange(W).float() + 0.5, indexing="ij")

    def image():
        cx = W / 2 + (torch.rand(B, 1, 1, generator=g) - 0.5) * W * 0.3
        cy = H / 2 + (torch.rand(B, 1, 1, generator=g) - 0.5) * H * 0.3
        rad = (0.18 + 0.22 * torch.rand(B, 1, 1, generator=g)) * min(H, W)
        dx, dy = (xx - cx) / rad, (yy - cy) / rad
        rr = dx * dx + dy * dy
        mask = (rr < 1).float().unsqueeze(1)                                  # [B,1,H,W]
        nz = torch.sqrt((1 - rr).clamp_min(0))
        normal = torch.stack([dx, dy, -nz], 1) * mask                         # [B,3,H,W]
        rgb = torch.rand(B, 3, H, W, generator=g) * mask + (1 - mask)
        ridx = torch.stack([torch.randperm(H * W, generator=g)[:R] for _ in range(B)])     # [B,R]

        def take(m):                                                          # [B,C,H,W] -> [B,R,C]
            flat = m.reshape(B, m.shape[1], H * W).permute(0, 2, 1)
            return torch.gather(flat, 1, ridx.unsqueeze(-1).expand(-1, -1, m.shape[1])).contiguous()
        return rgb, mask, normal, ridx, take(rgb), take(mask), take(normal)

    def pose_gt():
        az = torch.rand(B, generator=g) * 2 * math.pi
        z, o = torch.zeros(B), torch.ones(B)
        Rm = torch.stack([torch.stack([az.cos(), z, az.sin()], -1), torch.stack([z, o, z], -1),
                          torch.stack([-az.sin(), z, az.cos()], -1)], 1)
        return torch.cat([Rm, torch.tensor([0., 0., 5.]).expand(B, 3)[..., None]], -1)

    v = edict()
    v.idx = torch.arange(B)
    (v.rgb_input_map, v.mask_input_map, v.normal_input_map, v.ray_idx, v.rgb_input, v.mask_input, v.normal_input) = image()
    v.category_label = torch.zeros(B, dtype=torch.long)
    v.pose_gt = pose_gt()
    f = float(opt.camera.focal)
    v.intr = torch.tensor([[f * W, 0, W / 2], [0, f * H, H / 2], [0, 0, 1.]]).repeat(B, 1, 1)
    v.dpc = edict(points=torch.rand(B, 256, 3, generator=g) - 0.5)
    cols = {k: [] for k in ("rgb_input_map", "mask_input_map", "normal_input_map", "ray_idx", "rgb_input", "mask_input",
                            "normal_input", "pose_gt")}
    for _ in range(K):
        for k, t in zip(cols, image() + (pose_gt(),)):
            cols[k].append(t)
    for k, lst in cols.items():
        v[k + "_NN"] = torch.stack(lst, dim=-1).contiguous()
    for k in list(v.keys()):
        if isinstance(v[k], torch.Tensor):
            v[k] = v[k].to(device)
    v.dpc.points = v.dpc.points.to(device)
    return v


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["ref", "shim"], required=True)
    ap.add_argument("--case", choices=["train", "eval"], default="train")
    ap.add_argument("--out", required=True)
    ap.add_argument("--batch", type=int, default=3)
    ap.add_argument("--image", type=int, default=64)
    ap.add_argument("--rays", type=int, default=256)
    a = ap.parse_args()

    import numpy as np
    import torch
    import refharness as rh
    device = "cuda:0" if a.mode == "shim" else "cpu"
    if a.mode == "shim":
        from shapeclipper_b200 import shim
        shim.install()                                    # BEFORE the reference imports model.renderer / model.implicit / chamfer_3D
    mods = rh.import_reference_graph()
    import importlib
    runner = importlib.import_module("model.runner")
    if a.mode == "shim":
        assert mods.graph.Renderer.__module__.startswith("shapeclipper_b200"), "shim not active"
        assert mods.graph.SDFNetwork.__module__.startswith("shapeclipper_b200"), "shim not active"
    # the out-of-scope CNNs (cuDNN) feed the renderer: keep them in full fp32 so the comparison with the CPU run is about the hot path
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    edict = mods.util.EasyDict
    opt = rh.load_reference_opt(H=a.image, W=a.image)
    opt.image_size = [a.image, a.image]
    opt.render.rand_sample = a.rays
    opt.arch.enc_pretrained = False
    opt.device = device

    torch.manual_seed(0)
    np.random.seed(0)
    graph = mods.graph.Graph(opt)                         # constructed on the CPU in both modes: same seeded initial weights
    with torch.no_grad():                                 # move the SDF net off its latent-independent sphere initialisation
        gen = torch.Generator().manual_seed(7)
        for p in list(graph.sdf_network.parameters()) + list(graph.rgb_network.parameters()):
            p.add_(0.004 * torch.randn(p.shape, generator=gen))
    graph = graph.to(device)
    graph.eval()                                          # BatchNorm in eval mode, as Runner does while it <= iter_camera
    out = {"state": {k: v.detach().cpu().clone() for k, v in graph.state_dict().items()
                     if k.startswith(("sdf_network", "rgb_network", "renderer"))}}
    var = make_var(edict, opt, a.batch, seed=11, device=device)

    if a.case == "train":
        torch.manual_seed(1)
        np.random.seed(1)
        var, loss = graph.forward(opt, var, training=True, get_loss=True)
        loss = runner.Runner.summarize_loss(None, opt, var, loss)
        loss.all.backward()
        out["loss"] = {k: float(v) for k, v in loss.items()}
        out["idx_NN"] = var.input_NN_0.ray_idx.detach().cpu()
        for k in ("rgb_recon", "mask_recon", "mask_hard", "depth_recon", "normal_recon", "grad_eikonal", "rgb_recon_NN_0",
                  "mask_recon_NN_0", "normal_recon_NN_0", "proj_latent_sdf", "proj_latent_rgb", "pose", "scale_dist"):
            out[k] = var[k].detach().cpu()
        out["grad"] = {n: p.grad.detach().cpu() for n, p in graph.named_parameters()
                       if p.grad is not None and n.startswith(("sdf_network", "rgb_network", "renderer", "latent_proj",
                                                               "estimator.extr_fc", "estimator.size_fc", "encoder.fc"))}
    else:
        with torch.no_grad():
            opt.H, opt.W = 16, 16                         # Runner.evaluate_batch: full-grid render at the eval size
            for k in ("rgb", "mask", "normal"):           # evaluation batches carry whole maps, not ray samples
                m = var[k + "_input_map"]
                var[k + "_input"] = m.reshape(m.shape[0], m.shape[1], -1).permute(0, 2, 1).contiguous()
            torch.manual_seed(1)
            var = graph.forward(opt, var, training=False, get_loss=False)
            for k in ("rgb_recon_map", "mask_recon_map", "mask_hard_map", "normal_recon_map", "depth_recon"):
                out[k] = var[k].detach().cpu()
            eval_3D = importlib.import_module("utils.eval_3D")
            opt.eval.vox_res = 24
            pts = eval_3D.get_dense_3D_grid(opt, var)
            out["level"] = eval_3D.compute_level_grid(opt, graph.sdf_network, var.proj_latent_sdf, pts).cpu()
    torch.save(out, a.out)
    print("dropin_worker %s/%s ok" % (a.mode, a.case))


if __name__ == "__main__":
    main()
