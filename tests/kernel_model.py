"""Test-side executable model of the ALGORITHM the CUDA render kernels implement (not the oracle, not
product code): folded per-image latent biases, reverse-mode d(sdf)/dx ("gradient pass"), closed-form
compositing adjoints and the hand-derived double-backward. tests/test_kernel_model.py checks it in fp64
against autograd over oracle/render_ref.py, so that the derivation is pinned before it is written in CUDA.
Variable names match shapeclipper_b200/csrc/render_*.cu.
"""
import math

import torch

N_PE = 39
FREQS = [1.0, 2.0, 4.0, 8.0, 16.0, 32.0]
INV_SQRT2 = 1.0 / math.sqrt(2.0)


def fold_sdf(P):
    """nn.Linear weights -> the per-point matrices and the latent matrices (1/sqrt2 folded in)."""
    W0, W1, W2 = P["lin0.weight"], P["lin1.weight"], P["lin2.weight"]
    d = dict(A0=W0[:, :39], Z0=W0[:, 39:], b0=P["lin0.bias"],
             B1=W1[:, :64] * INV_SQRT2, A1=W1[:, 64:103] * INV_SQRT2, Z1=W1[:, 103:] * INV_SQRT2, b1=P["lin1.bias"],
             B2=W2[:, :64] * INV_SQRT2, A2=W2[:, 64:103] * INV_SQRT2, Z2=W2[:, 103:] * INV_SQRT2, b2=P["lin2.bias"],
             W3=P["lin3.weight"], b3=P["lin3.bias"], W4=P["lin4.weight"], b4=P["lin4.bias"],
             w5=P["lin5.weight"][0], b5=P["lin5.bias"][0], W5f=P["lin5.weight"][1:], b5f=P["lin5.bias"][1:])
    return d


def fold_rgb(P):
    V0 = P["lin0.weight"]
    return dict(V0p=V0[:, :39], V0z=V0[:, 39:103], V0f=V0[:, 103:], c0=P["lin0.bias"],
                V1=P["lin1.weight"], c1=P["lin1.bias"], V2=P["lin2.weight"], c2=P["lin2.bias"],
                V3=P["lin3.weight"], c3=P["lin3.bias"])


def posenc_all(xt):
    """xt [P,3] (already mirrored) -> pe [P,39], dpe (d pe_k / d xt_c(k)) [P,39], d2pe [P,39]."""
    pe, d1, d2 = [xt], [torch.ones_like(xt)], [torch.zeros_like(xt)]
    for f in FREQS:
        s, c = torch.sin(xt * f), torch.cos(xt * f)
        pe += [s, c]
        d1 += [f * c, -f * s]
        d2 += [-f * f * s, -f * f * c]
    return torch.cat(pe, -1), torch.cat(d1, -1), torch.cat(d2, -1)


def coord_reduce(v):
    """[P,39] per-feature values -> [P,3] summed over the features of each coordinate (feature k -> k % 3)."""
    return v.view(v.shape[0], 13, 3).sum(1)


def coord_expand(v3):
    """[P,3] -> [P,39] (each feature gets its coordinate's value)."""
    return v3.repeat(1, 13)


def softplus100(a):
    t = 100.0 * a
    e = torch.exp(torch.clamp(t, max=20.0))
    h = torch.where(t > 20.0, a, torch.log1p(e) / 100.0)
    s = torch.where(t > 20.0, torch.ones_like(a), e / (1.0 + e))
    tt = torch.where(t > 20.0, torch.zeros_like(a), 100.0 * s * (1.0 - s))
    return h, s, tt


def sdf_forward(F, x, cb):
    """x [P,3]; cb = dict(c0,c1,c2) per-point [P,64] latent biases. Returns the full stash."""
    sgn = torch.where(x[:, :1] < 0, -torch.ones_like(x[:, :1]), torch.ones_like(x[:, :1]))
    S = torch.cat([sgn, torch.ones_like(x[:, 1:])], -1)
    xt = torch.cat([x[:, :1].abs(), x[:, 1:]], -1)
    pe, dpe, d2pe = posenc_all(xt)
    st = dict(S=S, pe=pe, dpe=dpe, d2pe=d2pe)
    st["h0"], st["s0"], st["t0"] = softplus100(pe @ F["A0"].T + cb["c0"])
    st["h1"], st["s1"], st["t1"] = softplus100(st["h0"] @ F["B1"].T + pe @ F["A1"].T + cb["c1"])
    st["h2"], st["s2"], st["t2"] = softplus100(st["h1"] @ F["B2"].T + pe @ F["A2"].T + cb["c2"])
    st["h3"], st["s3"], st["t3"] = softplus100(st["h2"] @ F["W3"].T + F["b3"])
    st["h4"], st["s4"], st["t4"] = softplus100(st["h3"] @ F["W4"].T + F["b4"])
    st["sdf"] = st["h4"] @ F["w5"] + F["b5"]
    st["feat"] = st["h4"] @ F["W5f"].T + F["b5f"]
    return st


def sdf_gradient_pass(F, st):
    """reverse mode for the scalar sdf: gx = d sdf / d x."""
    st["q4"] = F["w5"].expand_as(st["s4"])
    st["g4"] = st["q4"] * st["s4"]
    st["q3"] = st["g4"] @ F["W4"];  st["g3"] = st["q3"] * st["s3"]
    st["q2"] = st["g3"] @ F["W3"];  st["g2"] = st["q2"] * st["s2"]
    st["q1"] = st["g2"] @ F["B2"];  st["g1"] = st["q1"] * st["s1"]
    st["q0"] = st["g1"] @ F["B1"];  st["g0"] = st["q0"] * st["s0"]
    st["gpe"] = st["g0"] @ F["A0"] + st["g1"] @ F["A1"] + st["g2"] @ F["A2"]
    st["gxt"] = coord_reduce(st["gpe"] * st["dpe"])
    st["gx"] = st["gxt"] * st["S"]
    return st


def sdf_backward(F, st, sdf_bar, feat_bar, gx_bar, pe_bar_extra=None):
    """Adjoint of sdf_forward + sdf_gradient_pass. sdf_bar [P], feat_bar [P,64] or None, gx_bar [P,3] or None,
    pe_bar_extra [P,39] (from the RGB net's own posenc input). Returns dict of weight grads (folded layout),
    per-point c-bias grads (c0b,c1b,c2b [P,64]) and x_bar [P,3]."""
    P = st["pe"].shape[0]
    G = {}
    z64 = torch.zeros_like(st["h0"])
    gb = {l: z64.clone() for l in range(5)}      # adjoint of g_l
    sb = {l: z64.clone() for l in range(5)}      # adjoint of s_l
    xt_bar = torch.zeros_like(st["S"])
    pe_bar = torch.zeros_like(st["pe"]) if pe_bar_extra is None else pe_bar_extra.clone()
    for k in ("A0", "A1", "A2", "B1", "B2", "W3", "W4"):
        G[k] = torch.zeros_like(F[k])
    G["w5"] = torch.zeros_like(F["w5"])
    if gx_bar is not None:
        gxt_bar = gx_bar * st["S"]
        gpe_bar = coord_expand(gxt_bar) * st["dpe"]                    # = J gxt_bar
        xt_bar = xt_bar + gxt_bar * coord_reduce(st["d2pe"] * st["gpe"])
        for l, A in ((0, "A0"), (1, "A1"), (2, "A2")):
            gb[l] = gb[l] + gpe_bar @ F[A].T
            G[A] = G[A] + st["g%d" % l].T @ gpe_bar
        qb0 = gb[0] * st["s0"]; sb[0] = sb[0] + gb[0] * st["q0"]
        gb[1] = gb[1] + qb0 @ F["B1"].T; G["B1"] = G["B1"] + st["g1"].T @ qb0
        qb1 = gb[1] * st["s1"]; sb[1] = sb[1] + gb[1] * st["q1"]
        gb[2] = gb[2] + qb1 @ F["B2"].T; G["B2"] = G["B2"] + st["g2"].T @ qb1
        qb2 = gb[2] * st["s2"]; sb[2] = sb[2] + gb[2] * st["q2"]
        gb[3] = qb2 @ F["W3"].T; G["W3"] = G["W3"] + st["g3"].T @ qb2
        qb3 = gb[3] * st["s3"]; sb[3] = gb[3] * st["q3"]
        gb[4] = qb3 @ F["W4"].T; G["W4"] = G["W4"] + st["g4"].T @ qb3
        G["w5"] = G["w5"] + (gb[4] * st["s4"]).sum(0); sb[4] = gb[4] * st["q4"]
    # first-order part
    hb4 = sdf_bar[:, None] * F["w5"][None]
    G["w5"] = G["w5"] + (sdf_bar[:, None] * st["h4"]).sum(0)
    G["b5"] = sdf_bar.sum()
    if feat_bar is not None:
        hb4 = hb4 + feat_bar @ F["W5f"]
        G["W5f"] = feat_bar.T @ st["h4"]; G["b5f"] = feat_bar.sum(0)
    else:
        G["W5f"] = torch.zeros_like(F["W5f"]); G["b5f"] = torch.zeros_like(F["b5f"])
    ab4 = hb4 * st["s4"] + sb[4] * st["t4"]
    G["W4"] = G["W4"] + ab4.T @ st["h3"]; G["b4"] = ab4.sum(0)
    ab3 = (ab4 @ F["W4"]) * st["s3"] + sb[3] * st["t3"]
    G["W3"] = G["W3"] + ab3.T @ st["h2"]; G["b3"] = ab3.sum(0)
    ab2 = (ab3 @ F["W3"]) * st["s2"] + sb[2] * st["t2"]
    G["B2"] = G["B2"] + ab2.T @ st["h1"]; G["A2"] = G["A2"] + ab2.T @ st["pe"]; pe_bar = pe_bar + ab2 @ F["A2"]
    ab1 = (ab2 @ F["B2"]) * st["s1"] + sb[1] * st["t1"]
    G["B1"] = G["B1"] + ab1.T @ st["h0"]; G["A1"] = G["A1"] + ab1.T @ st["pe"]; pe_bar = pe_bar + ab1 @ F["A1"]
    ab0 = (ab1 @ F["B1"]) * st["s0"] + sb[0] * st["t0"]
    G["A0"] = G["A0"] + ab0.T @ st["pe"]; pe_bar = pe_bar + ab0 @ F["A0"]
    xt_bar = xt_bar + coord_reduce(pe_bar * st["dpe"])
    return dict(G=G, c0b=ab0, c1b=ab1, c2b=ab2, x_bar=xt_bar * st["S"])


def unfold_sdf_grads(G, cb_sums, z_sdf_per_image):
    """Folded grads -> nn.Linear-layout grads. cb_sums = dict(c0,c1,c2) each [B,64] (per-image sums of the
    c-bias adjoints); returns (param grads dict, z_sdf grad [B,64] contribution)."""
    raise NotImplementedError  # done in the test with explicit algebra


# ----------------------------------------------------------------------------- compositing

def density_terms(sdf, beta):
    """sigma, cfac = -dsigma/dsdf (>0), and their derivatives."""
    e = 0.5 * torch.exp(-sdf.abs() / beta)
    sigma = torch.where(sdf >= 0, e, 1 - e) / beta
    cfac = e / (beta * beta)                                   # (1/(2 beta^2)) exp(-|s|/beta)
    sg = torch.where(sdf >= 0, torch.ones_like(sdf), -torch.ones_like(sdf))
    dsigma_dbeta = -sigma / beta + e * sdf / beta ** 3
    dc_ds = -sg / beta * cfac
    dc_dbeta = cfac * (-2.0 / beta + sdf.abs() / beta ** 2)
    return sigma, cfac, dsigma_dbeta, dc_ds, dc_dbeta


def safe_normalize(u, eps=1e-12):
    n = u.norm(dim=-1, keepdim=True)
    return u / n.clamp_min(eps), n


def safe_normalize_bwd(u, out, nrm, out_bar, eps=1e-12):
    big = nrm > eps
    return torch.where(big, (out_bar - out * (out * out_bar).sum(-1, keepdim=True)) / nrm.clamp_min(eps), out_bar / eps)


def composite_forward(z, sdf, gx, color, depth_fac, beta, bg=1.0, normal_pow=1.0):
    """z,sdf [N,S]; gx,color [N,S,3]; depth_fac [N]."""
    sigma, cfac, *_ = density_terms(sdf, beta)
    delta = torch.cat([z[:, 1:] - z[:, :-1], torch.zeros_like(z[:, :1])], -1)
    E = delta * sigma
    C = torch.cumsum(E, -1) - E                                # exclusive prefix sum
    T = torch.exp(-C)
    ea = torch.exp(-E)
    w = (1 - ea) * T
    u = cfac.unsqueeze(-1) * gx
    n_s, n_norm = safe_normalize(u)
    wp = w if normal_pow == 1.0 else w ** normal_pow
    Nsum = (wp.unsqueeze(-1) * n_s).sum(1)
    normal, Nn = safe_normalize(Nsum)
    acc = w.sum(-1)
    rgb = (w.unsqueeze(-1) * color).sum(1) + (1 - acc).unsqueeze(-1) * bg
    depth = (w * z).sum(-1) * depth_fac
    return dict(sigma=sigma, cfac=cfac, delta=delta, E=E, T=T, ea=ea, w=w, u=u, n_s=n_s, n_norm=n_norm, Nsum=Nsum,
                Nn=Nn, normal=normal, acc=acc, rgb=rgb, depth=depth, wp=wp)


def composite_backward(z, sdf, gx, color, depth_fac, beta, cf, rgb_bar, mask_bar, depth_bar, normal_bar,
                       bg=1.0, normal_pow=1.0):
    """-> sdf_bar [N,S], gx_bar [N,S,3], color_bar [N,S,3], z_bar [N,S], depth_fac_bar [N], beta_bar scalar."""
    sigma, cfac, dsig_dbeta, dc_ds, dc_dbeta = density_terms(sdf, beta)
    w, T, ea, delta = cf["w"], cf["T"], cf["ea"], cf["delta"]
    w_bar = (rgb_bar.unsqueeze(1) * (color - bg)).sum(-1) + mask_bar.unsqueeze(1) \
        + depth_bar.unsqueeze(1) * z * depth_fac.unsqueeze(1)
    color_bar = w.unsqueeze(-1) * rgb_bar.unsqueeze(1)
    z_bar = depth_bar.unsqueeze(1) * w * depth_fac.unsqueeze(1)
    depth_fac_bar = depth_bar * (w * z).sum(-1)
    Nsum_bar = safe_normalize_bwd(cf["Nsum"], cf["normal"], cf["Nn"], normal_bar)
    if normal_pow == 1.0:
        w_bar = w_bar + (Nsum_bar.unsqueeze(1) * cf["n_s"]).sum(-1)
    else:
        w_bar = w_bar + normal_pow * w ** (normal_pow - 1) * (Nsum_bar.unsqueeze(1) * cf["n_s"]).sum(-1)
    ns_bar = cf["wp"].unsqueeze(-1) * Nsum_bar.unsqueeze(1)
    u_bar = safe_normalize_bwd(cf["u"], cf["n_s"], cf["n_norm"], ns_bar)
    gx_bar = cfac.unsqueeze(-1) * u_bar
    c_bar = (u_bar * gx).sum(-1)
    alpha_bar = w_bar * T
    T_bar = w_bar * (1 - ea)
    C_bar = -T_bar * T
    suffix = torch.flip(torch.cumsum(torch.flip(C_bar, [-1]), -1), [-1]) - C_bar     # sum_{i>j} C_bar_i
    E_bar = suffix + alpha_bar * ea
    sigma_bar = E_bar * delta
    delta_bar = E_bar * sigma
    z_bar = z_bar - delta_bar                                   # delta_{S-1} is a constant 0: no gradient
    z_bar[:, -1] += delta_bar[:, -1]
    z_bar[:, 1:] += delta_bar[:, :-1]
    sdf_bar = sigma_bar * (-cfac) + c_bar * dc_ds
    beta_bar = (sigma_bar * dsig_dbeta).sum() + (c_bar * dc_dbeta).sum()
    return sdf_bar, gx_bar, color_bar, z_bar, depth_fac_bar, beta_bar
