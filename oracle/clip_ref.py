"""ORACLE (test infrastructure): torch restatement of the CLIP ViT image tower + the cosine k-NN of the annotator.

PARITY UNPINNED against openai/CLIP itself: the reference depends on `pip install git+https://github.com/openai/CLIP`
(README.md:14,24 — unpinned, not vendored, not in requirements.yaml, absent from /root/reference and from this image;
no network). Call sites: CLIP_anno.py:16 (clip.load("ViT-L/14")), :166 (encode_image(...).float()), :167 (F.normalize),
:29-57 (calc_matches). The reference holds no test or golden vector for this boundary. What this file restates is the
published architecture of openai/CLIP's VisionTransformer:

    x = conv1(img)  (kernel = stride = patch, no bias)  -> [B, W, G, G] -> [B, G*G, W]
    x = cat([class_embedding, x]) + positional_embedding ; x = ln_pre(x)
    for each block: x = x + attn(ln_1(x)) ; x = x + c_proj(QuickGELU(c_fc(ln_2(x))))      (nn.MultiheadAttention, d_head 64)
    x = ln_post(x[:, 0]) @ proj                                                            QuickGELU(x) = x * sigmoid(1.702 x)

and it is pinned against the one independent implementation available offline, HuggingFace transformers'
CLIPVisionModelWithProjection (an architecture twin: tests/test_oracle_clip.py maps the weights and compares).
Parameter names follow openai/CLIP's state_dict under `visual.` so real checkpoints map 1:1.
"""
import torch
import torch.nn.functional as F

CONFIGS = {
    "ViT-B/32": dict(image_size=224, patch=32, width=768, layers=12, heads=12, out_dim=512),
    "ViT-L/14": dict(image_size=224, patch=14, width=1024, layers=24, heads=16, out_dim=768),
    "tiny": dict(image_size=64, patch=32, width=128, layers=2, heads=2, out_dim=64),
}


def random_params(cfg, seed=0, dtype=torch.float32):
    """Random-init parameters with openai/CLIP's shapes and init scales."""
    g = torch.Generator().manual_seed(seed)
    W, P, L = cfg["width"], cfg["patch"], cfg["layers"]
    T = (cfg["image_size"] // P) ** 2 + 1
    s = W ** -0.5
    rn = lambda *shape, std=1.0: (torch.randn(*shape, generator=g) * std).to(dtype)
    p = {"conv1.weight": rn(W, 3, P, P, std=(3 * P * P) ** -0.5), "class_embedding": rn(W, std=s),
         "positional_embedding": rn(T, W, std=s), "ln_pre.weight": 1 + rn(W, std=0.02), "ln_pre.bias": rn(W, std=0.02),
         "ln_post.weight": 1 + rn(W, std=0.02), "ln_post.bias": rn(W, std=0.02), "proj": rn(W, cfg["out_dim"], std=s)}
    for i in range(L):
        b = "transformer.resblocks.%d." % i
        p[b + "ln_1.weight"], p[b + "ln_1.bias"] = 1 + rn(W, std=0.02), rn(W, std=0.02)
        p[b + "attn.in_proj_weight"], p[b + "attn.in_proj_bias"] = rn(3 * W, W, std=s), rn(3 * W, std=0.02)
        p[b + "attn.out_proj.weight"], p[b + "attn.out_proj.bias"] = rn(W, W, std=s * (2 * L) ** -0.5), rn(W, std=0.02)
        p[b + "ln_2.weight"], p[b + "ln_2.bias"] = 1 + rn(W, std=0.02), rn(W, std=0.02)
        p[b + "mlp.c_fc.weight"], p[b + "mlp.c_fc.bias"] = rn(4 * W, W, std=(2 * W) ** -0.5), rn(4 * W, std=0.02)
        p[b + "mlp.c_proj.weight"], p[b + "mlp.c_proj.bias"] = rn(W, 4 * W, std=s * (2 * L) ** -0.5), rn(W, std=0.02)
    return p


def encode_image(p, cfg, images):
    """images [B,3,S,S] (already normalised) -> [B, out_dim] (not L2-normalised, like clip's encode_image)."""
    W, H = cfg["width"], cfg["heads"]
    x = F.conv2d(images, p["conv1.weight"], stride=cfg["patch"])
    B = x.shape[0]
    x = x.reshape(B, W, -1).permute(0, 2, 1)
    x = torch.cat([p["class_embedding"].expand(B, 1, W), x], dim=1) + p["positional_embedding"]
    x = F.layer_norm(x, (W,), p["ln_pre.weight"], p["ln_pre.bias"], 1e-5)
    T = x.shape[1]
    for i in range(cfg["layers"]):
        b = "transformer.resblocks.%d." % i
        y = F.layer_norm(x, (W,), p[b + "ln_1.weight"], p[b + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, p[b + "attn.in_proj_weight"], p[b + "attn.in_proj_bias"])
        q, k, v = [t.reshape(B, T, H, W // H).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
        att = torch.softmax((q * (W // H) ** -0.5) @ k.transpose(-1, -2), dim=-1) @ v
        x = x + F.linear(att.transpose(1, 2).reshape(B, T, W), p[b + "attn.out_proj.weight"], p[b + "attn.out_proj.bias"])
        y = F.layer_norm(x, (W,), p[b + "ln_2.weight"], p[b + "ln_2.bias"], 1e-5)
        h = F.linear(y, p[b + "mlp.c_fc.weight"], p[b + "mlp.c_fc.bias"])
        x = x + F.linear(h * torch.sigmoid(1.702 * h), p[b + "mlp.c_proj.weight"], p[b + "mlp.c_proj.bias"])
    x = F.layer_norm(x[:, 0], (W,), p["ln_post.weight"], p["ln_post.bias"], 1e-5)
    return x @ p["proj"]


def calc_matches(features, k_nearest=6, thres=None):
    """NN_annotator.calc_matches (CLIP_anno.py:29-57): per query cosine to all + top-k; with `thres` (opt.thres) the
    query followed by k-1 random picks among the similarities in [thres, 1) when there are enough of them."""
    idx, val = [], []
    for i in range(features.shape[0]):
        cos = (features[i:i + 1] * features).sum(dim=1)
        if thres is not None:
            index = ((cos >= thres) & (cos < 1.)).nonzero()
            n_valid = len(index)
            if n_valid >= k_nearest - 1:
                sampled = index[torch.randperm(n_valid)[:k_nearest - 1]].squeeze(1)
                both = torch.cat([torch.tensor([i]), sampled], dim=0)
                idx.append(both); val.append(cos[both])
                continue
        v, j = cos.topk(k_nearest, largest=True)
        idx.append(j); val.append(v)
    return torch.stack(idx), torch.stack(val)
