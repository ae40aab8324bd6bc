"""Host side of the single-kernel CLIP image tower (csrc/clip_tower.cu; C ABI sc_clip_tower_* in include/sc_b200.h).

Weight preparation ("folding", once per model):
  * LayerNorm gains go into the matrix that consumes the normalised activations, W'[n,k] = g[k] W[n,k]; the kernel then
    evaluates LN(x) W^T + b as rstd (x W'^T - mean s) + c with s[n] = sum_k W'[n,k] (of the ROUNDED 16-bit planes, so the mean
    term cancels exactly what the MMA accumulated) and c[n] = sum_k ln_bias[k] W[n,k] + b[n];
  * the 1/sqrt(64) of the attention scores goes into the q rows of W'_qkv and c_qkv (exact: a power of two);
  * matrices become 16-bit planes: fp16 (`precision="fp16"`, one MMA per product — what `clip.load` itself runs on CUDA,
    CLIP_anno.py:16) or hi/lo bf16 pairs (`precision="split"`, three MMAs per product, fp32-class).
A plan (phase table + every TMA descriptor, device resident) is built once per batch size and reused by every encode.
"""
import ctypes

import torch

from . import _lib

_vp = ctypes.c_void_p


class ScClipTowerLayer(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("qkv_w_hi", "qkv_w_lo", "qkv_s", "qkv_c", "out_w_hi", "out_w_lo", "out_b",
                                   "fc1_w_hi", "fc1_w_lo", "fc1_s", "fc1_c", "fc2_w_hi", "fc2_w_lo", "fc2_b")]


class ScClipTowerWeights(ctypes.Structure):
    _fields_ = [(n, _vp) for n in ("conv_w_hi", "conv_w_lo", "class_emb", "pos_emb", "lnpre_w", "lnpre_b", "lnpost_w",
                                   "lnpost_b", "proj_t")] + [("layers", ctypes.POINTER(ScClipTowerLayer))]


def declare(L, cfg_type):
    i, sz = ctypes.c_int, ctypes.c_size_t
    L.sc_clip_tower_workspace_bytes.argtypes = [ctypes.POINTER(cfg_type), i]
    L.sc_clip_tower_workspace_bytes.restype = sz
    L.sc_clip_tower_plan_bytes.argtypes = [ctypes.POINTER(cfg_type)]
    L.sc_clip_tower_plan_bytes.restype = sz
    L.sc_clip_tower_plan.argtypes = [ctypes.POINTER(cfg_type), ctypes.POINTER(ScClipTowerWeights), i, _vp, sz, _vp, sz,
                                     ctypes.POINTER(ctypes.c_int), _vp]
    L.sc_clip_tower_plan.restype = i
    L.sc_clip_tower_workspace_layout.argtypes = [ctypes.POINTER(cfg_type), i, ctypes.POINTER(sz * 16)]
    L.sc_clip_tower_workspace_layout.restype = i
    L.sc_clip_tower_encode.argtypes = [ctypes.POINTER(cfg_type), _vp, i, _vp, _vp, _vp, _vp, _vp, _vp, i, _vp]
    L.sc_clip_tower_encode.restype = i


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _planes(w, split):
    """fp32 matrix -> (hi, lo | None, row sums of the rounded planes in fp32)."""
    w = w.detach().float().contiguous()
    if split:
        hi = w.to(torch.bfloat16)
        lo = (w - hi.float()).to(torch.bfloat16)
        s = (hi.double() + lo.double()).sum(1).float()
        return hi.contiguous(), lo.contiguous(), s.contiguous()
    hi = w.to(torch.float16)
    return hi.contiguous(), None, hi.double().sum(1).float().contiguous()


class TowerWeights:
    """Folded device copies of a CLIPVisual's parameters + the ctypes structs that point at them."""

    def __init__(self, params, cfg, split):
        P_ = params
        W, Ln = cfg["width"], cfg["layers"]
        self.keep = []
        self.split = bool(split)

        def vec(t):
            v = t.detach().float().contiguous()
            self.keep.append(v)
            return _p(v)

        def mat(w):
            hi, lo, s = _planes(w, split)
            self.keep.extend([hi, lo, s])
            return _p(hi), _p(lo), s
        self.layers = (ScClipTowerLayer * Ln)()
        for l in range(Ln):
            b = "transformer.resblocks.%d." % l
            Lr = self.layers[l]
            g1, b1 = P_[b + "ln_1.weight"].detach().float(), P_[b + "ln_1.bias"].detach().float()
            g2, b2 = P_[b + "ln_2.weight"].detach().float(), P_[b + "ln_2.bias"].detach().float()
            wqkv, bqkv = P_[b + "attn.in_proj_weight"].detach().float(), P_[b + "attn.in_proj_bias"].detach().float()
            scale = torch.ones(3 * W, device=wqkv.device)
            scale[:W] = 0.125                                     # 1 / sqrt(head dim 64) on the q rows
            Lr.qkv_w_hi, Lr.qkv_w_lo, s = mat(wqkv * g1[None, :] * scale[:, None])
            Lr.qkv_s = vec(s)
            Lr.qkv_c = vec(((wqkv.double() @ b1.double()) + bqkv.double()).float() * scale)
            Lr.out_w_hi, Lr.out_w_lo, _ = mat(P_[b + "attn.out_proj.weight"])
            Lr.out_b = vec(P_[b + "attn.out_proj.bias"])
            w1, bb1 = P_[b + "mlp.c_fc.weight"].detach().float(), P_[b + "mlp.c_fc.bias"].detach().float()
            Lr.fc1_w_hi, Lr.fc1_w_lo, s = mat(w1 * g2[None, :])
            Lr.fc1_s = vec(s)
            Lr.fc1_c = vec(((w1.double() @ b2.double()) + bb1.double()).float())
            Lr.fc2_w_hi, Lr.fc2_w_lo, _ = mat(P_[b + "mlp.c_proj.weight"])
            Lr.fc2_b = vec(P_[b + "mlp.c_proj.bias"])
        w = ScClipTowerWeights()
        conv = P_["conv1.weight"].detach().float().reshape(W, -1)
        kpad = (-conv.shape[1]) % 64                              # ViT-L/14: 3*14*14 = 588 -> 640 (the im2col phase pads alike)
        if kpad:
            conv = torch.cat([conv, torch.zeros(W, kpad, device=conv.device)], dim=1)
        w.conv_w_hi, w.conv_w_lo, _ = mat(conv)
        w.class_emb, w.pos_emb = vec(P_["class_embedding"]), vec(P_["positional_embedding"])
        w.lnpre_w, w.lnpre_b = vec(P_["ln_pre.weight"]), vec(P_["ln_pre.bias"])
        w.lnpost_w, w.lnpost_b = vec(P_["ln_post.weight"]), vec(P_["ln_post.bias"])
        w.proj_t = vec(P_["proj"].detach().float().t())
        w.layers = ctypes.cast(self.layers, ctypes.POINTER(ScClipTowerLayer))
        self.struct = w


class TowerPlan:
    """Workspace + device-resident phase table / TMA descriptors for one (model, batch size)."""

    def __init__(self, cfg_struct, weights, batch, device):
        L = _lib.lib()
        self.batch = batch
        with torch.cuda.device(device):
            ws_bytes = L.sc_clip_tower_workspace_bytes(ctypes.byref(cfg_struct), batch)
            plan_bytes = L.sc_clip_tower_plan_bytes(ctypes.byref(cfg_struct))
            self.workspace = torch.zeros(ws_bytes, dtype=torch.uint8, device=device)
            self.plan = torch.zeros(plan_bytes, dtype=torch.uint8, device=device)
            n = ctypes.c_int(0)
            stream = torch.cuda.current_stream(device)
            _lib.check(L.sc_clip_tower_plan(ctypes.byref(cfg_struct), ctypes.byref(weights.struct), batch, _p(self.workspace), ws_bytes,
                                            _p(self.plan), plan_bytes, ctypes.byref(n), ctypes.c_void_p(stream.cuda_stream)),
                       "sc_clip_tower_plan")
            self.n_phases = n.value


LAYOUT_NAMES = ("x", "x16_hi", "x16_lo", "qkv_hi", "qkv_lo", "attn_hi", "attn_lo", "h_hi", "h_lo", "patch_hi", "patch_lo", "patch_out",
                "y", "raw", "stats_a", "stats_b")


def workspace_layout(cfg_struct, batch):
    """{buffer name: byte offset inside TowerPlan.workspace} (diagnostics / tests)."""
    arr = (ctypes.c_size_t * 16)()
    _lib.check(_lib.lib().sc_clip_tower_workspace_layout(ctypes.byref(cfg_struct), batch, ctypes.byref(arr)), "sc_clip_tower_workspace_layout")
    return dict(zip(LAYOUT_NAMES, [int(v) for v in arr]))


def encode(cfg_struct, plan, images, emb, raw, hi, lo, per_phase_launches=False, stop_after=None, group=0):
    """One tower encode on the current stream of images' device (1 memset + 1 cooperative launch; per_phase_launches=True
    runs the same device code as one ordinary launch per phase — what ncu attributes GEMM by GEMM; group = g > 0: cooperative
    launches of g phases each, for a tower that shares the GPU with another stream's kernels). Returns the number of launches."""
    L = _lib.lib()
    with torch.cuda.device(images.device):
        _lib.check(L.sc_clip_tower_encode(ctypes.byref(cfg_struct), _p(plan.plan), plan.n_phases, _p(plan.workspace), _p(images),
                                          _p(emb), _p(raw), _p(hi), _p(lo),
                                          (stop_after + 1) if stop_after else (1 if per_phase_launches else -int(group)),
                                          _lib.stream_of(images)), "sc_clip_tower_encode")
    if per_phase_launches:
        return plan.n_phases
    pieces = 1 if group <= 0 else 1 + -(-(plan.n_phases - 3) // int(group))       # the three embedding phases, then `group` at a time
    return 2 * pieces                                                              # a counter memset + a cooperative launch each
