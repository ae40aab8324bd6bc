"""GPU: the drop-in claim. The reference's OWN code — model/graph.py Graph.forward (68-265), model/runner.py
Runner.summarize_loss (294-305), utils/eval_3D.py (9-49,105-121,155-165) — runs with `shapeclipper_b200.shim.install()`
underneath on CUDA and must reproduce what the same code computes unshimmed on the CPU from the same seeds.

The reference modules come from tests/refharness.py (/root/reference in the build container; on the GPU box the sourceless
bytecode that oracle/build_ref.py staged under oracle/_ref/py). The two runs of the training step live in separate
processes (tests/dropin_worker.py) so the two `model.renderer` modules never share an interpreter.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import refharness

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
needs_ref = pytest.mark.skipif(not refharness.reference_available(), reason="reference modules not staged (run __graft_entry__.build() in the build container)")


def _worker(mode, case, out):
    r = subprocess.run([sys.executable, os.path.join(HERE, "dropin_worker.py"), "--mode", mode, "--case", case, "--out", out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return torch.load(out)


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@needs_ref
def test_reference_graph_train_step_under_shim(tmp_path):
    """Graph.forward(training=True, get_loss=True) + Runner.summarize_loss + backward: every loss, every render output and
    the gradients of the hot-path parameters AND of the CNN-side layers that feed it."""
    ref = _worker("ref", "train", str(tmp_path / "ref.pt"))
    got = _worker("shim", "train", str(tmp_path / "shim.pt"))
    # same seeded construction: our SDFNetwork / RGBNetwork / Renderer consume the generator exactly as the reference's do
    assert ref["state"].keys() == got["state"].keys()
    for k in ref["state"]:
        assert torch.equal(ref["state"][k], got["state"][k]), k
    assert torch.equal(ref["idx_NN"], got["idx_NN"])                   # same np.random.choice neighbour per sample
    report = {}
    for k in ("proj_latent_sdf", "proj_latent_rgb", "pose", "scale_dist"):     # cuDNN vs CPU convolutions upstream of the renderer
        report["cnn." + k] = _rel(got[k], ref[k])
        assert report["cnn." + k] < 5e-5, (k, report["cnn." + k])
    for k in ("rgb_recon", "mask_recon", "depth_recon", "rgb_recon_NN_0", "mask_recon_NN_0"):
        err = float((got[k] - ref[k]).abs().max())
        report[k] = err
        assert err < 1e-4, (k, err)
    flips = float((got["mask_hard"] != ref["mask_hard"]).float().mean())
    assert flips <= 2e-3, flips
    for k in ("normal_recon", "normal_recon_NN_0"):                    # unit normals, weighted by the mask they matter under
        w = ref["mask_recon" if k == "normal_recon" else "mask_recon_NN_0"]
        err = float(((got[k] - ref[k]).abs() * w).max())
        report[k] = err
        assert err < 2e-4, (k, err)
    err = float((got["grad_eikonal"] - ref["grad_eikonal"]).abs().max())
    report["grad_eikonal"] = err
    assert err < 3e-4, err
    # losses: 1e-4 relative (north_star); the two trimmed normal losses select pixels by a hard threshold on the predicted mask
    # and by rank, so one pixel of ~200 changing sides moves them by more than the arithmetic does: 2e-3
    for k, v in ref["loss"].items():
        tol = 2e-3 if "normal" in k else 1e-4
        assert abs(got["loss"][k] - v) <= tol * max(abs(v), 1e-3), (k, got["loss"][k], v)
        report["loss." + k] = abs(got["loss"][k] - v) / max(abs(v), 1e-3)
    assert ref["grad"].keys() == got["grad"].keys()
    worst = ("", 0.0)
    for n, g in ref["grad"].items():
        r = _rel(got["grad"][n], g)
        if r > worst[1]:
            worst = (n, r)
        # fp32 double backward through a 6-layer softplus(100) MLP: the reference's own CPU fp32 gradients sit 1e-4..1e-3
        # (relative to the largest entry) away from an fp64 evaluation (tests/test_render_gpu.py), so 3e-3 here
        assert r < 3e-3, (n, r)
    report["worst_grad"] = worst
    print("drop-in train step:", report)


@needs_ref
def test_reference_graph_eval_and_level_grid_under_shim(tmp_path):
    """Runner.evaluate_batch's call: Graph.forward(training=False, get_loss=False) — full-grid render reshaped to maps — and
    utils/eval_3D.get_dense_3D_grid + compute_level_grid (slice by slice) through the shimmed SDFNetwork."""
    ref = _worker("ref", "eval", str(tmp_path / "ref.pt"))
    got = _worker("shim", "eval", str(tmp_path / "shim.pt"))
    for k in ("rgb_recon_map", "mask_recon_map", "depth_recon"):
        err = float((got[k] - ref[k]).abs().max())
        assert err < 1e-4, (k, err)
    assert float((got["mask_hard_map"] != ref["mask_hard_map"]).float().mean()) <= 2e-3
    err = float(((got["normal_recon_map"] - ref["normal_recon_map"]).abs() * ref["mask_recon_map"]).max())
    assert err < 2e-4, err
    assert got["level"].shape == ref["level"].shape == (3, 25, 25, 25)
    err = float((got["level"] - ref["level"]).abs().max())
    assert err < 2e-5, err


@needs_ref
def test_reference_eval3d_functions_under_shim():
    """utils/eval_3D.chamfer_distance / compute_fscore / normalize_pc of the reference, calling OUR chamfer_3D module through
    the shim, against this repo's eval_3D.py (bit-equal) and the C oracle (bit-equal distances, equal indices)."""
    import importlib
    from oracle import chamfer_ref
    from shapeclipper_b200 import eval_3D as ours, shim
    shim.install()
    for name in ("mcubes", "trimesh"):
        refharness._stub(name)
    refharness.import_reference()
    ref_eval = importlib.import_module("utils.eval_3D")
    import chamfer_3D
    assert chamfer_3D.__name__ == "shapeclipper_b200.chamfer_3D" and ref_eval.chamfer_3D is chamfer_3D
    opt = refharness.load_reference_opt()
    opt.device = "cuda:0"
    g = torch.Generator().manual_seed(3)
    pred = (torch.randn(2, 3000, 3, generator=g) * 0.3).cuda()
    gt = (torch.randn(2, 2500, 3, generator=g) * 0.3 + 0.02).cuda()
    a, b = ref_eval.normalize_pc(pred), ref_eval.normalize_pc(gt)
    assert torch.equal(a, ours.normalize_pc(pred)) and torch.equal(b, ours.normalize_pc(gt))
    d1, d2, i1, i2 = ref_eval.chamfer_distance(opt, a, b)
    e1, e2, j1, j2 = ours.chamfer_distance(opt, a, b)
    assert torch.equal(d1, e1) and torch.equal(d2, e2) and torch.equal(i1, j1) and torch.equal(i2, j2)
    o1, o2, k1, k2 = chamfer_ref.chamfer_forward(a.cpu().numpy(), b.cpu().numpy())
    assert (i1.cpu().numpy() == k1).all() and (i2.cpu().numpy() == k2).all()
    assert np.array_equal(np.sqrt(o1).astype(np.float32).view(np.int32), d1.cpu().numpy().view(np.int32))
    assert np.array_equal(np.sqrt(o2).astype(np.float32).view(np.int32), d2.cpu().numpy().view(np.int32))
    f_ref = ref_eval.compute_fscore(d1, d2, opt.eval.f_thresholds)
    f_ours = ours.compute_fscore(d1, d2, opt.eval.f_thresholds)
    assert torch.equal(f_ref, f_ours) and f_ref.shape == (2, 6)


def test_hotpathgraph_step_matches_oracle_step():
    """HotPathGraph (forward + seven losses + backward, B=4, 512 rays x 64 samples, CPU-generator draws, neighbour 0 injected)
    against the CPU restatement of the same step (bench.oracle_step_fn's body): every loss and every parameter / leaf gradient."""
    from oracle import loss_ref, render_ref as R
    from shapeclipper_b200 import options, synthetic
    from shapeclipper_b200.graph import HotPathGraph
    from shapeclipper_b200.options import Options
    dev = torch.device("cuda:0")
    opt = options.default_options(device=str(dev))
    B = 4
    torch.manual_seed(0)
    graph = HotPathGraph(opt)
    with torch.no_grad():
        gen = torch.Generator().manual_seed(7)
        for p in list(graph.sdf_network.parameters()) + list(graph.rgb_network.parameters()):
            p.add_(0.004 * torch.randn(p.shape, generator=gen))
    batch = synthetic.make_batch(opt, B, seed=21, pin=False)
    # ---- oracle (CPU fp32 autograd)
    sp = {k: v.detach().clone().requires_grad_(True) for k, v in graph.sdf_network.state_dict().items()}
    rp = {k: v.detach().clone().requires_grad_(True) for k, v in graph.rgb_network.state_dict().items()}
    beta = graph.renderer.density.beta.detach().clone().requires_grad_(True)
    leaves = {k: batch[k].clone().requires_grad_(True) for k in ("pose", "intr", "scale_dist", "proj_latent_sdf", "proj_latent_rgb")}
    torch.manual_seed(5)
    out = R.render(sp, rp, beta, leaves["pose"], leaves["intr"], leaves["scale_dist"], leaves["proj_latent_sdf"],
                   leaves["proj_latent_rgb"], opt.H, opt.W, ray_idx=batch["ray_idx"], training=True)
    L = loss_ref.render_losses(out, batch["rgb_input"], batch["mask_input"], batch["normal_input"] @ leaves["pose"][..., :3], B)
    out2 = R.render(sp, rp, beta, batch["pose_NN"][..., 0], batch["intr_NN"][..., 0], batch["scale_dist_NN"][..., 0],
                    leaves["proj_latent_sdf"], batch["proj_latent_rgb_NN"][..., 0], opt.H, opt.W,
                    ray_idx=batch["ray_idx_NN"][..., 0], training=True)
    L2 = loss_ref.render_losses(out2, batch["rgb_input_NN"][..., 0], batch["mask_input_NN"][..., 0],
                                batch["normal_input_NN"][..., 0] @ batch["pose_NN"][..., 0][..., :3], B)
    want = dict(L, nearest_img=L2["render"], nearest_mask=L2["mask"], nearest_normal=L2["normal"])
    total = loss_ref.weighted_total(L) + 1.0 * L2["render"] + 0.5 * L2["mask"] + 0.01 * L2["normal"]
    total.backward()
    # ---- ours
    graph = graph.to(dev)
    graph.select_neighbours = lambda o, v: torch.zeros(B, 1, dtype=torch.long, device=dev)     # injected draw: neighbour 0
    var = Options()
    for k, t in batch.items():
        var[k] = t.to(dev)
        if k in leaves:
            var[k].requires_grad_(True)
    torch.manual_seed(5)
    _, loss = graph(opt, var, training=True, get_loss=True)
    loss["all"].backward()
    report = {}
    for k, v in want.items():
        tol = 2e-3 if "normal" in k else 1e-4
        report[k] = abs(float(loss[k]) - float(v)) / max(abs(float(v)), 1e-3)
        assert report[k] <= tol, (k, float(loss[k]), float(v))
    assert abs(float(loss["all"]) - float(total)) <= 1e-4 * abs(float(total))
    worst = ("", 0.0)
    pairs = [("sdf." + k, graph.sdf_network.get_parameter(k).grad, v.grad) for k, v in sp.items()]
    pairs += [("rgb." + k, graph.rgb_network.get_parameter(k).grad, v.grad) for k, v in rp.items()]
    pairs += [("beta", graph.renderer.density.beta.grad, beta.grad)]
    pairs += [(k, var[k].grad, v.grad) for k, v in leaves.items()]
    for n, g, w in pairs:
        assert g is not None and w is not None, n
        r = _rel(g.cpu(), w)
        if r > worst[1]:
            worst = (n, r)
        assert r < 3e-3, (n, r)
    print("HotPathGraph vs oracle step:", report, "worst grad", worst)
