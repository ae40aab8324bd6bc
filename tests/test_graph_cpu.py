"""CPU: host logic of the graph / evaluation layer against the reference's own functions — live when the reference modules
are importable (tests/refharness.py) and through fixtures they produced (tests/golden/{neighbours,eval3d,calc_matches}.pt,
generator tests/gen_golden.py --round2)."""
import os

import numpy as np
import pytest
import torch

import refharness
from shapeclipper_b200.options import Options, default_options


def _fx(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _select(fx, n_views):
    """HotPathGraph.select_neighbours (numpy path) on the fixture's masks under the fixture's numpy seed."""
    from shapeclipper_b200.graph import HotPathGraph
    opt = default_options(device="cpu")
    opt.reg.n_views = n_views
    opt.reg.sample_temp = fx["sample_temp"]
    var = Options(mask_input=fx["mask_input"], mask_input_NN=fx["mask_input_NN"])
    np.random.seed(fx["np_seed"])
    return HotPathGraph.select_neighbours(None, opt, var)


@pytest.mark.parametrize("n_views", [1, 2])
def test_select_neighbours_matches_reference_golden(golden_dir, n_views):
    """G2 (model/graph.py:119-142): IoU -> (1 - IoU)^T -> L1-normalise -> np.random.choice without replacement."""
    fx = _fx(golden_dir, "neighbours")["views%d" % n_views]
    got = _select(fx, n_views)
    assert got.dtype == torch.long and torch.equal(got, fx["choice"])
    assert len(set(fx["choice"].flatten().tolist())) >= 3          # the fixture is not degenerate


@pytest.mark.skipif(not refharness.reference_available(), reason="reference modules absent")
def test_select_neighbours_matches_live_reference():
    import gen_golden
    gm = refharness.import_reference_graph()
    for seed, n_views in ((21, 1), (22, 3)):
        opt = refharness.load_reference_opt(H=2, W=2)
        opt.reg.n_views = n_views
        var = gen_golden.neighbour_var(gm, opt, 9, 50, seed)
        np.random.seed(seed)
        want = gen_golden.reference_forward_NN_choice(gm, opt, var)
        got = _select(dict(mask_input=var.mask_input, mask_input_NN=var.mask_input_NN, np_seed=seed,
                           sample_temp=opt.reg.sample_temp), n_views)
        assert torch.equal(got, want)


def test_fscore_and_normalize_pc_match_reference_golden(golden_dir):
    """C3 (utils/eval_3D.py:40-49,105-121): product functions and the oracle restatement, bit-equal to the reference's output
    (same torch ops on the same CPU), including the precision + recall = 0 -> NaN -> 0 row."""
    from oracle import render_ref as R
    from shapeclipper_b200 import eval_3D
    fx = _fx(golden_dir, "eval3d")
    assert torch.equal(eval_3D.normalize_pc(fx["pc"]), fx["pc_normalized"])
    assert torch.equal(R.normalize_pc(fx["pc"]), fx["pc_normalized"])
    f = eval_3D.compute_fscore(fx["dist1"], fx["dist2"], fx["thresholds"])
    assert torch.equal(f, fx["fscore"]) and float(fx["fscore"][3].abs().max()) == 0.0
    assert torch.equal(R.fscore(fx["dist1"], fx["dist2"], fx["thresholds"]), fx["fscore"])


@pytest.mark.skipif(not refharness.reference_available(), reason="reference modules absent")
def test_fscore_and_normalize_pc_match_live_reference():
    import importlib
    from shapeclipper_b200 import eval_3D
    for name in ("mcubes", "trimesh", "chamfer_3D"):
        refharness._stub(name)
    refharness.import_reference()
    ev = importlib.import_module("utils.eval_3D")
    g = torch.Generator().manual_seed(4)
    pc = torch.randn(2, 333, 3, generator=g)
    assert torch.equal(eval_3D.normalize_pc(pc), ev.normalize_pc(pc))
    d1, d2 = torch.rand(3, 100, generator=g) * 0.3, torch.rand(3, 120, generator=g) * 0.3
    assert torch.equal(eval_3D.compute_fscore(d1, d2), ev.compute_fscore(d1, d2))


def test_oracle_calc_matches_matches_reference_golden(golden_dir):
    """L2 (CLIP_anno.py:29-57): the oracle restatement against NN_annotator.calc_matches itself — top-k branch and both
    sub-branches of the opt.thres branch, with the reference's CPU-generator randperm draws."""
    from oracle import clip_ref
    fx = _fx(golden_dir, "calc_matches")
    idx, val = clip_ref.calc_matches(fx["features"], 6)
    assert torch.equal(idx, fx["topk"]["indices"]) and torch.equal(val, fx["topk"]["values"])
    for key in ("thres_0.55", "thres_0.895"):
        c = fx[key]
        torch.manual_seed(c["seed"])
        idx, val = clip_ref.calc_matches(fx["features"], 6, thres=c["thres"])
        assert torch.equal(idx, c["indices"]) and torch.equal(val, c["values"]), key
    fell_back = (fx["thres_0.895"]["indices"] == fx["topk"]["indices"]).all(1)
    assert 0 < int(fell_back.sum()) < len(fell_back)                # both sub-branches are in the fixture


def test_neighbour_losses_are_gated_individually():
    """model/graph.py:241-263 gates nearest_img / nearest_mask / nearest_normal each on its own weight."""
    from shapeclipper_b200.graph import HotPathGraph
    opt = default_options(device="cpu")
    g = HotPathGraph(opt)
    gen = torch.Generator().manual_seed(0)
    B, R = 2, 40
    r = lambda *s: torch.rand(*s, generator=gen)
    unit = lambda t: torch.nn.functional.normalize(t - 0.5, dim=-1)
    var = Options(rgb_recon=r(B, R, 3), mask_recon=r(B, R, 1), normal_recon=unit(r(B, R, 3)), grad_eikonal=r(B * 2 * R),
                  rgb_input=r(B, R, 3), mask_input=(r(B, R, 1) > 0.3).float(), normal_transformed=unit(r(B, R, 3)),
                  rgb_recon_NN_0=r(B, R, 3), mask_recon_NN_0=r(B, R, 1), normal_recon_NN_0=unit(r(B, R, 3)),
                  pose_NN_0=torch.eye(3, 4).repeat(B, 1, 1))
    var["input_NN_0"] = dict(rgb_input=r(B, R, 3), mask_input=(r(B, R, 1) > 0.3).float(), normal_input=unit(r(B, R, 3)))
    full = g.compute_loss(opt, var, training=True)
    assert {"nearest_img", "nearest_mask", "nearest_normal"} <= set(full)
    opt.loss_weight.nearest_img = None
    opt.loss_weight.nearest_normal = None
    only_mask = g.compute_loss(opt, var, training=True)
    assert "nearest_mask" in only_mask and "nearest_img" not in only_mask and "nearest_normal" not in only_mask
    assert torch.equal(only_mask["nearest_mask"], full["nearest_mask"])
    want = sum(float(getattr(opt.loss_weight, k)) * only_mask[k] for k in ("render", "mask", "normal", "eikonal", "nearest_mask"))
    assert torch.allclose(only_mask["all"], want)
