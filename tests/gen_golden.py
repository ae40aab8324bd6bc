"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference,
build container only). Run:  python tests/gen_golden.py
The fixtures pin both the oracle (oracle/render_ref.py) and the CUDA path; they carry every input,
the reference's CPU-generator draws, its outputs and — for the training case — its autograd gradients.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import refharness as rh  # noqa: E402
from oracle import render_ref as R  # noqa: E402

OUT = os.path.join(HERE, "golden")


def make_networks(mods, opt, seed, perturb=0.02):
    torch.manual_seed(seed)
    sdf = mods.implicit.SDFNetwork(opt)
    rgb = mods.implicit.RGBNetwork(opt)
    ren = mods.renderer.Renderer(opt, sdf, rgb)
    with torch.no_grad():   # geometric init zeroes the latent columns: perturb so every weight matters
        for p in sdf.parameters():
            p.add_(perturb * torch.randn_like(p))
    return sdf, rgb, ren


def random_camera(mods, opt, B):
    az = torch.rand(B) * 6.2831853
    el = (torch.rand(B) - 0.5)
    th = (torch.rand(B) - 0.5) * 0.3

    def trig(a):
        return torch.stack([a.cos(), a.sin()], -1)
    Ry = mods.camera.azim_to_rotation_matrix(trig(az), "trig")
    Rx = mods.camera.elev_to_rotation_matrix(trig(el), "trig")
    Rz = mods.camera.roll_to_rotation_matrix(trig(th), "trig")
    P = torch.tensor([[-1, 0, 0], [0, 0, -1], [0, -1, 0]]).float()
    Rm = Rz @ Rx @ Ry @ P
    sd = 1 + 0.1 * (torch.rand(B) - 0.5)
    t = torch.stack([torch.zeros(B), torch.zeros(B), sd * opt.camera.dist], -1)
    pose = torch.cat([Rm, t[..., None]], -1)
    intr = mods.camera.get_intr(opt, 1 + 0.05 * (torch.rand(B) - 0.5))
    return pose, intr, sd


def case_render(mods, name, H, W, B, n_rays, training, seed):
    opt = rh.load_reference_opt(H=H, W=W)
    sdf, rgb, ren = make_networks(mods, opt, seed)
    pose, intr, sd = random_camera(mods, opt, B)
    zs, zr = torch.randn(B, 64) * 0.3, torch.randn(B, 64) * 0.3
    ray_idx = None
    if n_rays is not None:
        ray_idx = torch.stack([torch.randperm(H * W)[:n_rays] for _ in range(B)])
    Rn = n_rays if n_rays is not None else H * W
    S = opt.render.n_samples_uniform
    leaves = dict(pose=pose, intr=intr, scale_dist=sd, z_sdf=zs, z_rgb=zr)
    for v in leaves.values():
        v.requires_grad_(training)
    torch.manual_seed(seed + 100)
    state = torch.get_rng_state()
    outs = ren(opt, pose, intr, sd, zs, zr, ray_idx=ray_idx, training=training)
    torch.set_rng_state(state)
    u, eik_idx, eik_pts = R.draw_render_rng(B * Rn, S, training)
    names = ["rgb", "mask", "mask_hard", "depth", "normal", "grad_eik"]
    fx = dict(H=H, W=W, B=B, S=S, training=training, ray_idx=ray_idx,
              sdf_params={k: v.detach().clone() for k, v in sdf.state_dict().items()},
              rgb_params={k: v.detach().clone() for k, v in rgb.state_dict().items()},
              beta=ren.density.beta.detach().clone(),
              inputs={k: v.detach().clone() for k, v in leaves.items()},
              rng=dict(u=u, eik_idx=eik_idx, eik_pts=eik_pts),
              outputs={n: (o.detach().clone() if o is not None else None) for n, o in zip(names, outs)})
    if training:
        # fixed cotangents -> one scalar -> the reference's autograd gradients
        torch.manual_seed(seed + 200)
        cot = {n: torch.randn_like(o) for n, o in zip(names, outs) if o is not None and n != "mask_hard"}
        scalar = sum((cot[n] * o).sum() for n, o in zip(names, outs) if n in cot)
        params = dict(ren.named_parameters())   # density.beta + sdf_network.* + rgb_network.*
        wrt = list(params.values()) + list(leaves.values())
        grads = torch.autograd.grad(scalar, wrt, allow_unused=True)
        keys = list(params.keys()) + list(leaves.keys())
        fx["cotangents"] = cot
        fx["grads"] = {k: (g.detach().clone() if g is not None else None) for k, g in zip(keys, grads)}
    # fp64 evaluation of the same inputs/draws with the oracle restatement ("truth"): lets the GPU tests bound their
    # error by the reference's own fp32 rounding noise on ill-conditioned entries (grazing-ray normals, beta).
    dt = torch.float64
    c64 = lambda t: t.detach().to(dt) if t.is_floating_point() else t.detach()
    sp64 = {k: c64(v).requires_grad_(training) for k, v in sdf.state_dict().items()}
    rp64 = {k: c64(v).requires_grad_(training) for k, v in rgb.state_dict().items()}
    in64 = {k: c64(v).requires_grad_(training) for k, v in leaves.items()}
    beta64 = c64(ren.density.beta).requires_grad_(training)
    o64 = R.render(sp64, rp64, beta64, in64["pose"], in64["intr"], in64["scale_dist"], in64["z_sdf"], in64["z_rgb"],
                   H, W, ray_idx=ray_idx, training=training, rng=(u, eik_idx, eik_pts))
    fx["outputs64"] = {n: (o64[n].detach().clone() if o64[n] is not None else None) for n in names}
    if training:
        scalar64 = sum((cot[n].to(dt) * o64[n].view_as(cot[n])).sum() for n in cot)
        wrt64 = [beta64] + list(sp64.values()) + list(rp64.values()) + list(in64.values())
        keys64 = ["density.beta"] + ["sdf_network." + k for k in sp64] + ["rgb_network." + k for k in rp64] + list(in64.keys())
        g64 = torch.autograd.grad(scalar64, wrt64, allow_unused=True)
        fx["grads64"] = {k: (g.detach().clone() if g is not None else None) for k, g in zip(keys64, g64)}
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print("wrote", name, {n: (tuple(o.shape) if o is not None else None) for n, o in fx["outputs"].items()})


def case_visualize(mods, name, H, W, B, seed):
    """Renderer.forward(..., training=False, visualize=True): the three extra debug tensors (model/renderer.py:174-183)."""
    opt = rh.load_reference_opt(H=H, W=W)
    sdf, rgb, ren = make_networks(mods, opt, seed)
    pose, intr, sd = random_camera(mods, opt, B)
    zs, zr = torch.randn(B, 64) * 0.3, torch.randn(B, 64) * 0.3
    torch.manual_seed(seed + 100)
    with torch.no_grad():
        outs = ren(opt, pose, intr, sd, zs, zr, ray_idx=None, training=False, visualize=True)
    names = ["rgb", "mask", "mask_hard", "depth", "normal", "grad_eik", "points_sampled", "transparency_sampled", "rgb_sampled"]
    torch.save(dict(H=H, W=W, B=B, seed=seed + 100,
                    sdf_params={k: v.detach().clone() for k, v in sdf.state_dict().items()},
                    rgb_params={k: v.detach().clone() for k, v in rgb.state_dict().items()},
                    beta=ren.density.beta.detach().clone(),
                    inputs=dict(pose=pose, intr=intr, scale_dist=sd, z_sdf=zs, z_rgb=zr),
                    outputs={n: (o.detach().clone() if o is not None else None) for n, o in zip(names, outs)}),
               os.path.join(OUT, name + ".pt"))
    print("wrote", name, [tuple(o.shape) for o in outs[6:]])


def case_sdf_query(mods, name, seed):
    opt = rh.load_reference_opt()
    sdf, _, _ = make_networks(mods, opt, seed)
    B, N = 3, 50
    pts = (torch.rand(B * N, 3) - 0.5) * 1.6
    z = torch.randn(B, 64) * 0.3
    s, f, g = sdf.get_conditional_output(opt, B, pts.clone(), z, compute_grad=True)
    torch.save(dict(sdf_params={k: v.detach().clone() for k, v in sdf.state_dict().items()},
                    pts=pts, z_sdf=z, B=B, sdf=s.detach(), feat=f.detach(), grad=g.detach()),
               os.path.join(OUT, name + ".pt"))
    print("wrote", name)


def case_losses(mods, name, seed):
    opt = rh.load_reference_opt()
    L = mods.loss.Loss(opt)
    torch.manual_seed(seed)
    B, Rn = 2, 300
    rgb, rgb_gt = torch.rand(B, Rn, 3), torch.rand(B, Rn, 3)
    mask, mask_gt = torch.rand(B, Rn, 1), (torch.rand(B, Rn, 1) > 0.4).float()
    n = torch.nn.functional.normalize(torch.randn(B, Rn, 3), dim=-1)
    n_gt = torch.nn.functional.normalize(n + 0.3 * torch.randn(B, Rn, 3), dim=-1)
    eik = 1 + 0.1 * torch.randn(B * 2 * Rn)
    valid = (mask_gt > 0.5) & (mask > 0.5)
    out = dict(render=L.MSE_loss(rgb, rgb_gt), mask=L.mask_loss(mask, mask_gt),
               normal=L.normal_loss(n, n_gt, valid, tolerance=opt.reg.normal_tol),
               eikonal=L.MSE_loss(eik.view(B, -1), 1))
    torch.save(dict(rgb=rgb, rgb_gt=rgb_gt, mask=mask, mask_gt=mask_gt, normal=n, normal_gt=n_gt, grad_eik=eik,
                    B=B, losses=out), os.path.join(OUT, name + ".pt"))
    print("wrote", name, {k: float(v) for k, v in out.items()})


def reference_forward_NN_choice(mods, opt, var):
    """Runs the reference's Graph.forward_NN (model/graph.py:114-218) with the CNNs / renderer of `self` replaced by dummies
    and returns the neighbour it selected per sample (model/graph.py:119-142: IoU -> (1-IoU)^T -> np.random.choice)."""
    import types
    B = var.mask_input.shape[0]
    z6 = tuple(torch.zeros(B, 1) for _ in range(6))
    fake = types.SimpleNamespace(encoder=lambda x: torch.zeros(B, opt.arch.latent_dim_shape + opt.arch.latent_dim_rgb),
                                 latent_proj_rgb=lambda x: x[:, :64],
                                 pred_pose=lambda *a, **k: (torch.zeros(B, 3, 4), torch.zeros(B, 3, 3), torch.ones(B)),
                                 renderer=lambda *a, **k: z6)
    mods.graph.Graph.forward_NN(fake, opt, var)
    return torch.stack([var["input_NN_%d" % v].ray_idx[:, 0] for v in range(opt.reg.n_views)], dim=1)   # ray_idx_NN[b,:,k] == k


def neighbour_var(mods, opt, B, R, seed):
    g = torch.Generator().manual_seed(seed)
    K = opt.data.k_nearest
    v = mods.util.EasyDict()
    v.idx = torch.arange(B)
    q = (torch.rand(B, R, 1, generator=g) > 0.5).float()
    flip = torch.rand(B, R, 1, K, generator=g) < torch.linspace(0.25, 0.45, K).view(1, 1, 1, K)      # IoU spread over the K neighbours
    v.mask_input = q
    v.mask_input_NN = torch.where(flip, 1 - q.unsqueeze(-1), q.unsqueeze(-1)).contiguous()
    v.rgb_input, v.normal_input = torch.zeros(B, R, 3), torch.zeros(B, R, 3)
    v.rgb_input_NN, v.normal_input_NN = torch.zeros(B, R, 3, K), torch.zeros(B, R, 3, K)
    v.rgb_input_map, v.mask_input_map, v.normal_input_map = torch.zeros(B, 3, 2, 2), torch.zeros(B, 1, 2, 2), torch.zeros(B, 3, 2, 2)
    v.rgb_input_map_NN, v.mask_input_map_NN = torch.zeros(B, 3, 2, 2, K), torch.zeros(B, 1, 2, 2, K)
    v.normal_input_map_NN = torch.zeros(B, 3, 2, 2, K)
    v.pose_gt, v.pose_gt_NN = torch.zeros(B, 3, 4), torch.zeros(B, 3, 4, K)
    v.ray_idx = torch.zeros(B, R, dtype=torch.long)
    v.ray_idx_NN = torch.arange(K).view(1, 1, K).expand(B, R, K).contiguous()
    v.proj_latent_sdf = torch.zeros(B, 64)
    return v


def case_neighbours(mods, name, seed):
    import numpy as np
    gm = rh.import_reference_graph()
    out = {}
    for n_views in (1, 2):
        opt = rh.load_reference_opt(H=2, W=2)
        opt.reg.n_views = n_views
        var = neighbour_var(gm, opt, 16, 96, seed)
        np.random.seed(seed)
        choice = reference_forward_NN_choice(gm, opt, var)
        out["views%d" % n_views] = dict(mask_input=var.mask_input.clone(), mask_input_NN=var.mask_input_NN.clone(), np_seed=seed,
                                       choice=choice.clone(), sample_temp=opt.reg.sample_temp)
    torch.save(out, os.path.join(OUT, name + ".pt"))
    print("wrote", name, out["views1"]["choice"].flatten().tolist())


def case_eval3d(mods, name, seed):
    """utils/eval_3D.py normalize_pc (40-49) and compute_fscore (105-121, incl. the NaN -> 0 rule) of the reference."""
    import importlib
    ev = importlib.import_module("utils.eval_3D")
    g = torch.Generator().manual_seed(seed)
    pc = torch.randn(3, 500, 3, generator=g) * torch.tensor([0.3, 0.5, 0.2]) + torch.tensor([0.1, -0.2, 0.05])
    d1 = torch.rand(4, 700, generator=g) * 0.15
    d2 = torch.rand(4, 650, generator=g) * 0.25
    d1[3] += 1.0
    d2[3] += 1.0                              # nothing under any threshold: precision + recall = 0 -> NaN -> 0
    th = [0.005, 0.01, 0.02, 0.05, 0.1, 0.2]
    torch.save(dict(pc=pc, pc_normalized=ev.normalize_pc(pc), dist1=d1, dist2=d2, thresholds=th,
                    fscore=ev.compute_fscore(d1, d2, th)), os.path.join(OUT, name + ".pt"))
    print("wrote", name)


def case_calc_matches(mods, name, seed):
    """NN_annotator.calc_matches (CLIP_anno.py:29-57) of the reference: top-k branch and the opt.thres sampling branch (its
    torch.randperm draws come from the CPU generator, seeded here). `clip` is absent: stubbed, never called."""
    import importlib
    rh._stub("clip")
    anno = importlib.import_module("CLIP_anno")
    ann = anno.NN_annotator.__new__(anno.NN_annotator)
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn(6, 64, generator=g)
    feats = torch.nn.functional.normalize(centres[torch.arange(90) % 6] + 0.35 * torch.randn(90, 64, generator=g), dim=-1)
    # norm 0.999: the self-similarity (0.998) is then robustly inside [thres, 1) for every implementation — with unit norms the
    # reference's `cos_sim < 1.` test on the query itself is decided by fp32 rounding of sum(f_i^2)
    feats = feats * 0.999
    out = dict(features=feats)
    opt = mods.util.EasyDict(thres=None, device="cpu")
    ind, val = ann.calc_matches(opt, feats, k_nearest=6)
    out["topk"] = dict(indices=torch.stack(ind), values=val)
    for thres in (0.55, 0.895):
        opt = mods.util.EasyDict(thres=thres, device="cpu")
        torch.manual_seed(seed)
        ind, val = ann.calc_matches(opt, feats, k_nearest=6)
        out["thres_%g" % thres] = dict(indices=torch.stack(ind), values=val, seed=seed, thres=thres)
    torch.save(out, os.path.join(OUT, name + ".pt"))
    print("wrote", name)


def main():
    os.makedirs(OUT, exist_ok=True)
    mods = rh.import_reference()
    if "--round2" in sys.argv:
        case_neighbours(mods, "neighbours", seed=8)
        case_eval3d(mods, "eval3d", seed=9)
        case_calc_matches(mods, "calc_matches", seed=10)
        return
    if "--only-visualize" in sys.argv:
        case_visualize(mods, "render_visualize_10x10", 10, 10, 2, seed=6)
        return
    case_visualize(mods, "render_visualize_10x10", 10, 10, 2, seed=6)
    case_render(mods, "render_eval_12x12", 12, 12, 2, None, False, seed=1)
    case_render(mods, "render_train_40rays", 16, 16, 2, 40, True, seed=2)
    case_render(mods, "render_train_full_8x8", 8, 8, 1, None, True, seed=3)
    case_sdf_query(mods, "sdf_query", seed=4)
    case_losses(mods, "losses", seed=5)
    case_neighbours(mods, "neighbours", seed=8)
    case_eval3d(mods, "eval3d", seed=9)
    case_calc_matches(mods, "calc_matches", seed=10)


if __name__ == "__main__":
    main()
