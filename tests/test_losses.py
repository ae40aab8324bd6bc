"""CPU: host-side losses of the product and the oracle restatement against values produced by the reference's
model/loss.py (tests/golden/losses.pt)."""
import os

import torch

from oracle import loss_ref


def _fx(golden_dir):
    return torch.load(os.path.join(golden_dir, "losses.pt"), weights_only=False)


def test_oracle_losses_match_reference(golden_dir):
    fx = _fx(golden_dir)
    out = dict(rgb=fx["rgb"], mask=fx["mask"], normal=fx["normal"], grad_eik=fx["grad_eik"])
    L = loss_ref.render_losses(out, fx["rgb_gt"], fx["mask_gt"], fx["normal_gt"], fx["B"])
    for k, v in fx["losses"].items():
        assert torch.allclose(L[k], v, rtol=1e-6, atol=1e-7), k


def test_product_losses_match_reference(golden_dir):
    from shapeclipper_b200 import loss, options
    fx = _fx(golden_dir)
    opt = options.default_options()
    fns = loss.Loss(opt)
    out = dict(rgb=fx["rgb"], mask=fx["mask"], normal=fx["normal"], grad_eikonal=fx["grad_eik"])
    tgt = dict(rgb=fx["rgb_gt"], mask=fx["mask_gt"], normal=fx["normal_gt"])
    L = loss.render_losses(fns, opt, out, tgt)
    for k, v in fx["losses"].items():
        assert torch.allclose(L[k], v, rtol=1e-5, atol=1e-7), (k, L[k], v)


def test_trimmed_normal_loss_gradient_matches_oracle():
    from shapeclipper_b200 import loss, options
    torch.manual_seed(0)
    opt = options.default_options()
    fns = loss.Loss(opt)
    n = torch.nn.functional.normalize(torch.randn(2, 100, 3), dim=-1).requires_grad_(True)
    g = torch.nn.functional.normalize(torch.randn(2, 100, 3), dim=-1)
    valid = torch.rand(2, 100, 1) > 0.3
    a = fns.normal_loss(n, g, valid, tolerance=0.2)
    ga = torch.autograd.grad(a, n)[0]
    n2 = n.detach().clone().requires_grad_(True)
    b = loss_ref.trimmed_normal(n2, g, valid.squeeze(-1), 0.2, 5.0)
    gb = torch.autograd.grad(b, n2)[0]
    assert torch.allclose(a, b, rtol=1e-6) and torch.allclose(ga, gb, rtol=1e-5, atol=1e-8)
