"""CPU, fp64: the hand-derived adjoints of tests/kernel_model.py (what the CUDA kernels implement) against
autograd over the oracle restatement."""
import torch

import kernel_model as km
from oracle import render_ref as R


def _rand_params(seed, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    dims = [(64, 103), (64, 167), (64, 167), (64, 64), (64, 64), (65, 64)]
    sp = {}
    for l, (o, i) in enumerate(dims):
        sp["lin%d.weight" % l] = (torch.randn(o, i, generator=g) * (0.25 / i ** 0.5 * 4)).to(dtype)
        sp["lin%d.bias" % l] = (torch.randn(o, generator=g) * 0.05).to(dtype)
    rp = {}
    for l, (o, i) in enumerate([(64, 167), (64, 64), (64, 64), (3, 64)]):
        rp["lin%d.weight" % l] = (torch.randn(o, i, generator=g) / i ** 0.5).to(dtype)
        rp["lin%d.bias" % l] = (torch.randn(o, generator=g) * 0.05).to(dtype)
    return sp, rp


def test_sdf_double_backward_matches_autograd():
    torch.manual_seed(0)
    sp, _ = _rand_params(1)
    sp = {k: v.requires_grad_(True) for k, v in sp.items()}
    Pn = 37
    x = ((torch.rand(Pn, 3, dtype=torch.float64) - 0.5) * 1.6).requires_grad_(True)
    z = (torch.randn(1, 64, dtype=torch.float64) * 0.3).requires_grad_(True)
    lat = z.expand(Pn, 64)
    out = R.sdf_mlp(sp, x, lat)
    sdf, feat = out[:, 0], out[:, 1:]
    gx = torch.autograd.grad(sdf.sum(), x, create_graph=True)[0]
    sdf_bar, feat_bar, gx_bar = torch.randn(Pn, dtype=torch.float64), torch.randn(Pn, 64, dtype=torch.float64), \
        torch.randn(Pn, 3, dtype=torch.float64)
    scalar = (sdf * sdf_bar).sum() + (feat * feat_bar).sum() + (gx * gx_bar).sum()
    names = list(sp.keys())
    ref = torch.autograd.grad(scalar, [x, z] + [sp[k] for k in names])
    ref_x, ref_z, ref_w = ref[0], ref[1], dict(zip(names, ref[2:]))

    with torch.no_grad():
        F = km.fold_sdf({k: v.detach() for k, v in sp.items()})
        zz = z.detach()[0]
        cb = {"c0": (F["Z0"] @ zz + F["b0"]).expand(Pn, 64), "c1": (F["Z1"] @ zz + F["b1"]).expand(Pn, 64),
              "c2": (F["Z2"] @ zz + F["b2"]).expand(Pn, 64)}
        st = km.sdf_gradient_pass(F, km.sdf_forward(F, x.detach(), cb))
        assert torch.allclose(st["sdf"], sdf.detach(), atol=1e-12)
        assert torch.allclose(st["feat"], feat.detach(), atol=1e-12)
        assert torch.allclose(st["gx"], gx.detach(), atol=1e-11)
        bw = km.sdf_backward(F, st, sdf_bar, feat_bar, gx_bar)
        assert torch.allclose(bw["x_bar"], ref_x, atol=1e-9, rtol=1e-9)
        G = bw["G"]
        s2 = km.INV_SQRT2
        c0, c1, c2 = bw["c0b"].sum(0), bw["c1b"].sum(0), bw["c2b"].sum(0)
        z_bar = F["Z0"].T @ c0 + F["Z1"].T @ c1 + F["Z2"].T @ c2
        assert torch.allclose(z_bar, ref_z[0], atol=1e-9, rtol=1e-9)
        W0g = torch.cat([G["A0"], torch.outer(c0, zz)], 1)
        W1g = torch.cat([G["B1"], G["A1"], torch.outer(c1, zz)], 1) * s2
        W2g = torch.cat([G["B2"], G["A2"], torch.outer(c2, zz)], 1) * s2
        W5g = torch.cat([G["w5"][None], G["W5f"]], 0)
        for name, got in (("lin0.weight", W0g), ("lin1.weight", W1g), ("lin2.weight", W2g), ("lin3.weight", G["W3"]),
                          ("lin4.weight", G["W4"]), ("lin5.weight", W5g), ("lin0.bias", c0), ("lin1.bias", c1),
                          ("lin2.bias", c2), ("lin3.bias", G["b3"]), ("lin4.bias", G["b4"]),
                          ("lin5.bias", torch.cat([G["b5"][None], G["b5f"]]))):
            assert torch.allclose(got, ref_w[name], atol=1e-8, rtol=1e-8), name


def test_composite_adjoints_match_autograd():
    torch.manual_seed(1)
    N, S = 9, 16
    dt = torch.float64
    z = (4.3 + torch.sort(torch.rand(N, S, dtype=dt) * 1.4, dim=-1)[0]).requires_grad_(True)
    sdf = (torch.randn(N, S, dtype=dt) * 0.2).requires_grad_(True)
    gx = torch.randn(N, S, 3, dtype=dt).requires_grad_(True)
    color = torch.rand(N, S, 3, dtype=dt).requires_grad_(True)
    f = (0.9 + 0.1 * torch.rand(N, dtype=dt)).requires_grad_(True)
    beta = torch.tensor(0.1, dtype=dt, requires_grad=True)
    for p in (1.0, 1.7):
        # autograd reference written with the oracle's pieces
        sigma = R.laplace_density(sdf, beta)
        w, _ = R.composite(z, sigma)
        dsig = torch.autograd.grad(sigma.sum(), sdf, create_graph=True)[0]
        n_flat = -dsig.unsqueeze(-1) * gx
        n_s = torch.nn.functional.normalize(n_flat, dim=-1)
        normal = torch.nn.functional.normalize(((w.unsqueeze(-1) ** p) * n_s).sum(1), dim=-1)
        acc = w.sum(-1)
        rgb = (w.unsqueeze(-1) * color).sum(1) + (1 - acc).unsqueeze(-1)
        depth = (w * z * f[:, None]).sum(1)
        bars = [torch.randn_like(rgb), torch.randn_like(acc), torch.randn_like(depth), torch.randn_like(normal)]
        scalar = (rgb * bars[0]).sum() + (acc * bars[1]).sum() + (depth * bars[2]).sum() + (normal * bars[3]).sum()
        ref = torch.autograd.grad(scalar, [sdf, gx, color, z, f, beta])
        with torch.no_grad():
            cf = km.composite_forward(z, sdf, gx, color, f, beta, 1.0, p)
            assert torch.allclose(cf["rgb"], rgb) and torch.allclose(cf["normal"], normal)
            assert torch.allclose(cf["depth"], depth) and torch.allclose(cf["acc"], acc)
            got = km.composite_backward(z, sdf, gx, color, f, beta, cf, *bars, bg=1.0, normal_pow=p)
        for a, b, nm in zip(got, ref, ["sdf", "gx", "color", "z", "depth_fac", "beta"]):
            assert torch.allclose(a, b, atol=1e-9, rtol=1e-8), (nm, p, (a - b).abs().max())
