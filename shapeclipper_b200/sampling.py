"""Boundary-distance ray sampler on the device.

Reference: `utils/util.py:237-248` — `compute_sampling_prob(opt, mask, uniform_fac)` is called per image in the DataLoader
workers (`data/pix3d.py:234-239`, train split with `opt.render.rand_sample`): `vigra.filters.boundaryDistanceTransform` of the
binarised mask on the CPU, `prob = normalize(1 / (sdf_2D + uniform_fac), p=1)`, then `np.random.choice(H*W, rand_sample, p=prob,
replace=False)`. Here the transform runs on the GPU (`csrc/sampler.cu`, exact Euclidean, bit-equal to scipy's exact EDT):

* `boundary_distance(mask)`                    the transform itself, [H,W] or [B,H,W]
* `compute_sampling_prob(opt, mask, fac)`      the reference's function with the reference's signature and its numpy draw (drop-in)
* `sample_rays(masks, n, fac, generator)`      the batched, sync-free form: an exponential race on the CUDA generator
                                               (-log(u) / p, the n smallest keys = a draw without replacement with the same law)
* `vigra`                                      a stand-in module object exposing `filters.boundaryDistanceTransform(ndarray)` that
                                               `shim.install()` registers when the real vigra is absent, so that the reference's own
                                               `utils.util.compute_sampling_prob` runs unchanged on this kernel
There is no CPU fallback: masks are moved to the current CUDA device."""
import ctypes
import types

import numpy as np
import torch

from . import _lib
from ._lib import ptr as _p


def declare(L):
    vp, i, sz, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
    L.sc_boundary_distance_scratch_bytes.argtypes = [i, i, i]
    L.sc_boundary_distance_scratch_bytes.restype = sz
    L.sc_boundary_distance.argtypes = [vp, i, i, i, f, vp, vp, vp, f, vp, vp]
    L.sc_boundary_distance.restype = i


def _launch(mask, threshold, want_dist, uniforms, uniform_fac):
    if not mask.is_cuda:
        raise ValueError("shapeclipper_b200.sampling needs a CUDA tensor (no CPU path)")
    m = mask.detach().to(torch.float32).contiguous()
    B, H, W = m.shape
    L = _lib.lib()
    scratch = torch.empty(L.sc_boundary_distance_scratch_bytes(B, H, W), dtype=torch.uint8, device=m.device)
    dist = torch.empty_like(m) if want_dist else None
    keys = torch.empty_like(m) if uniforms is not None else None
    with torch.cuda.device(m.device):
        _lib.check(L.sc_boundary_distance(_p(m), B, H, W, float(threshold), _p(scratch), _p(dist), _p(uniforms), float(uniform_fac),
                                          _p(keys), _lib.stream_of(m)), "sc_boundary_distance")
    return dist, keys


def boundary_distance(mask, threshold=0.5):
    """Euclidean distance of every pixel to the nearest pixel of the other class (mask > threshold), minus 0.5 — what
    vigra.filters.boundaryDistanceTransform((mask > 0.5).float()) returns with its default InterpixelBoundary. [H,W] or [B,H,W]."""
    squeeze = mask.dim() == 2
    d, _ = _launch(mask[None] if squeeze else mask, threshold, True, None, 0.0)
    return d[0] if squeeze else d


def sample_rays(masks, n, uniform_fac=3.0, generator=None, threshold=0.5):
    """[B,H,W] masks -> [B,n] int64 pixel indices (row-major, as var.ray_idx), n per image without replacement with probability
    proportional to 1 / (boundary distance + uniform_fac): the law of utils/util.py:245-247, drawn on the CUDA generator as an
    exponential race (no host round trip; capturable)."""
    B, H, W = masks.shape
    if n > H * W:
        raise ValueError("cannot draw %d of %d pixels without replacement" % (n, H * W))
    u = torch.rand(B, H, W, device=masks.device, generator=generator)
    u = 1.0 - u                                            # (0, 1]
    _, keys = _launch(masks, threshold, False, u, uniform_fac)
    return torch.topk(keys.view(B, H * W), n, dim=1, largest=False, sorted=True).indices


def compute_sampling_prob(opt, mask, uniform_fac=3):
    """utils/util.py:237-248 with its signature and its draw: mask [H,W] -> LongTensor [opt.render.rand_sample] on the CPU,
    `np.random.choice` on numpy's global generator exactly as the reference calls it; only the distance transform moved to the GPU."""
    assert len(mask.shape) == 2
    h, w = mask.shape
    assert opt.H == h
    dev = mask.device if mask.is_cuda else torch.device("cuda", torch.cuda.current_device())
    sdf_2D = boundary_distance(mask.to(dev)).cpu()
    prob_vec = 1 / (sdf_2D + uniform_fac)
    prob_vec = torch.nn.functional.normalize(prob_vec.view(h * w), dim=-1, p=1).cpu().numpy()
    return torch.tensor(np.random.choice(h * w, opt.render.rand_sample, p=prob_vec, replace=False))


def _vigra_boundary_distance_transform(array, *args, **kwargs):
    if args or kwargs:
        raise NotImplementedError("only vigra.filters.boundaryDistanceTransform(array) with its defaults is provided")
    a = np.asarray(array)
    if a.ndim != 2:
        raise NotImplementedError("2-D label images only")
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    # vigra treats the array as a LABEL image: the reference passes (mask > 0.5).float(), i.e. labels 0.0 / 1.0
    return boundary_distance(t, threshold=0.5).cpu().numpy()


vigra = types.ModuleType("shapeclipper_b200.sampling.vigra")
vigra.__doc__ = "stand-in for the part of vigra the reference uses (utils/util.py:243)"
vigra.filters = types.ModuleType("shapeclipper_b200.sampling.vigra.filters")
vigra.filters.boundaryDistanceTransform = _vigra_boundary_distance_transform
