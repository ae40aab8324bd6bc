"""GPU diagnostics for csrc/clip_tower.cu: runs the tower phase by phase (per-phase launches, stop_after=k) and compares the
workspace buffers of the patch embedding and of layer 0 with a torch fp32 evaluation; then the whole encode (cooperative
launch) with the oracle.   python scripts/debug_clip_tower.py [tiny|ViT-B/32|ViT-L/14] [batch] [split|fp16]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import clip_ref  # noqa: E402
from shapeclipper_b200 import clip, clip_tower  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-9))


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    prec = sys.argv[3] if len(sys.argv) > 3 else "split"
    cfg = clip_ref.CONFIGS[name]
    p = clip_ref.random_params(cfg, seed=2)
    torch.manual_seed(1)
    img = torch.randn(B, 3, cfg["image_size"], cfg["image_size"])
    vis = clip.CLIPVisual(name, precision=prec)
    vis.load_params(p)
    vis = vis.cuda()
    c, plan, dev = vis._tower_state(B)
    lay = clip_tower.workspace_layout(c, B)
    W, P, S = cfg["width"], cfg["patch"], cfg["image_size"]
    G = S // P
    T = G * G + 1
    M, Mp = B * T, B * G * G
    Kc = 3 * P * P
    Kp = (Kc + 63) // 64 * 64
    dt16 = torch.bfloat16 if prec == "split" else torch.float16
    ws = plan.workspace

    def buf(nm, shape, dtype):
        n = 1
        for s_ in shape:
            n *= s_
        nbytes = n * torch.empty(0, dtype=dtype).element_size()
        return ws[lay[nm]:lay[nm] + nbytes].view(dtype).view(*shape)

    def planes(nm, shape):
        hi = buf(nm + "_hi", shape, dt16).float()
        return hi + buf(nm + "_lo", shape, dt16).float() if prec == "split" else hi

    def run(k):
        D = cfg["out_dim"]
        emb = torch.zeros(B, D, device="cuda"); raw = torch.zeros(B, D, device="cuda")
        clip_tower.encode(c, plan, img.cuda(), emb, raw, None, None, stop_after=k)
        torch.cuda.synchronize()
        return raw, emb

    g = {k: v.cuda().float() for k, v in p.items()}
    x_img = img.cuda()
    patches = F.unfold(x_img, kernel_size=P, stride=P).transpose(1, 2).reshape(Mp, Kc)          # column = c*P*P + py*P + px
    run(1)
    print("phase 0 im2col      ", rel(planes("patch", (Mp, Kp))[:, :Kc], patches))
    po = patches @ g["conv1.weight"].reshape(W, -1).t()
    run(2)
    print("phase 1 patch GEMM  ", rel(buf("patch_out", (Mp, W), torch.float32), po))
    tok = torch.cat([g["class_embedding"].expand(B, 1, W), po.view(B, G * G, W)], 1) + g["positional_embedding"]
    x0 = F.layer_norm(tok, (W,), g["ln_pre.weight"], g["ln_pre.bias"], 1e-5).reshape(M, W)
    run(3)
    print("phase 2 tokens x    ", rel(buf("x", (M, W), torch.float32), x0), " x16", rel(planes("x16", (M, W)), x0))
    b0 = "transformer.resblocks.0."
    y = F.layer_norm(x0, (W,), g[b0 + "ln_1.weight"], g[b0 + "ln_1.bias"], 1e-5)
    qkv = y @ g[b0 + "attn.in_proj_weight"].t() + g[b0 + "attn.in_proj_bias"]
    qkv_s = qkv.clone(); qkv_s[:, :W] *= 0.125
    run(4)
    print("phase 3 qkv         ", rel(planes("qkv", (M, 3 * W)), qkv_s))
    H = cfg["heads"]
    q, k, v = [t.reshape(B, T, H, 64).transpose(1, 2) for t in qkv.chunk(3, dim=-1)]
    att = (torch.softmax((q * 0.125) @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(M, W)
    run(5)
    print("phase 4 attention   ", rel(planes("attn", (M, W)), att))
    x1 = x0 + att @ g[b0 + "attn.out_proj.weight"].t() + g[b0 + "attn.out_proj.bias"]
    run(6)
    print("phase 5 out-proj x  ", rel(buf("x", (M, W), torch.float32), x1))
    y2 = F.layer_norm(x1, (W,), g[b0 + "ln_2.weight"], g[b0 + "ln_2.bias"], 1e-5)
    h = y2 @ g[b0 + "mlp.c_fc.weight"].t() + g[b0 + "mlp.c_fc.bias"]
    h = h * torch.sigmoid(1.702 * h)
    run(7)
    print("phase 6 fc1 h       ", rel(planes("h", (M, 4 * W)), h))
    x2 = x1 + h @ g[b0 + "mlp.c_proj.weight"].t() + g[b0 + "mlp.c_proj.bias"]
    run(8)
    print("phase 7 fc2 x       ", rel(buf("x", (M, W), torch.float32), x2))
    with torch.no_grad():
        want = clip_ref.encode_image(p, cfg, img)
    vis.per_phase_launches = True
    raw, emb = vis.encode(img.cuda())
    print("whole tower, per-phase launches:", rel(raw.cpu(), want))
    vis.per_phase_launches = False
    raw2, emb2 = vis.encode(img.cuda())
    torch.cuda.synchronize()
    print("whole tower, one cooperative launch:", rel(raw2.cpu(), want), " bit-equal to per-phase:", bool(torch.equal(raw, raw2)))


if __name__ == "__main__":
    main()
