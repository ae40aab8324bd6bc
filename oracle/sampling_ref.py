"""ORACLE (test infrastructure, not the product): CPU restatement of the reference's boundary-distance ray sampler,
`utils/util.py:237-248` (`compute_sampling_prob`).

The distance transform itself lives in a third-party dependency that is absent from /root/reference and from this image:
`vigra.filters.boundaryDistanceTransform` (vigra 1.11.x, un-pinned conda dependency of the reference's environment). Its published
definition (vigra/multi_distance.hxx, `boundaryMultiDistance`, default `boundary = InterpixelBoundary`, `array_border_is_active = False`):
for every pixel the Euclidean distance to the nearest pixel carrying a DIFFERENT label, computed exactly (separable parabola
intersection), minus the 0.5 pixel the inter-pixel boundary lies in front of that pixel's centre.
PARITY UNPINNED against vigra itself. Pinned instead (tests/test_oracle_sampling.py):
  * `boundary_distance_bruteforce` — the definition, O(N^2), for small images;
  * `boundary_distance_scipy` — scipy.ndimage.distance_transform_edt (an independent exact Euclidean transform) at full size;
  * `compute_sampling_prob` below against the reference's own function run live with `vigra` stubbed by the scipy transform
    (same `np.random.seed` => same indices), frozen in tests/golden/ray_sampler.npz by tests/gen_golden_sampler.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module."""
import numpy as np


def boundary_distance_bruteforce(label):
    """label [H,W] bool -> float32 [H,W]; O((HW)^2): small images only. No pixel of the other class: H + W."""
    lab = np.asarray(label, dtype=bool)
    H, W = lab.shape
    ys, xs = np.mgrid[0:H, 0:W]
    pts = np.stack([ys.ravel(), xs.ravel()], 1).astype(np.int64)
    flat = lab.ravel()
    out = np.empty(H * W, dtype=np.float32)
    for i in range(H * W):
        other = pts[flat != flat[i]]
        if len(other) == 0:
            out[i] = float(H + W)
            continue
        d2 = ((other - pts[i]) ** 2).sum(1).min()
        out[i] = np.float32(np.sqrt(np.float32(d2))) - np.float32(0.5)
    return out.reshape(H, W)


def boundary_distance_scipy(label):
    """The same through scipy's exact Euclidean distance transform (distance to the nearest zero), once per class."""
    from scipy import ndimage
    lab = np.asarray(label, dtype=bool)
    H, W = lab.shape
    if lab.all() or not lab.any():
        return np.full((H, W), float(H + W), dtype=np.float32)
    d_in = ndimage.distance_transform_edt(lab)          # foreground pixels: distance to the nearest background pixel
    d_out = ndimage.distance_transform_edt(~lab)        # background pixels: distance to the nearest foreground pixel
    d = np.where(lab, d_in, d_out)
    # exact integers under the root: float32(sqrt(float64 n)) == sqrt(float32 n) (no double rounding for sqrt, 53 >= 2*24 + 2)
    d2 = np.rint(d * d).astype(np.int64)
    return (np.sqrt(d2.astype(np.float32)).astype(np.float32) - np.float32(0.5)).astype(np.float32)


def compute_sampling_prob(H, rand_sample, mask, uniform_fac=3, distance=boundary_distance_scipy):
    """utils/util.py:237-248 line by line (opt.H -> H, opt.render.rand_sample -> rand_sample); mask: torch tensor or ndarray [H,W]."""
    import torch
    mask = torch.as_tensor(mask)
    assert len(mask.shape) == 2
    h, w = mask.shape
    assert H == h
    mask_binary = (mask > 0.5)                                                       # util.py:241
    sdf_2D = distance(mask_binary.float().cpu().numpy() > 0.5)                       # util.py:243 (vigra on the CPU)
    sdf_2D = torch.from_numpy(np.asarray(sdf_2D, dtype=np.float32))                  # util.py:244
    prob_vec = 1 / (sdf_2D + uniform_fac)                                            # util.py:245
    prob_vec = torch.nn.functional.normalize(prob_vec.view(h * w), dim=-1, p=1).cpu().numpy()        # util.py:246
    indices = torch.tensor(np.random.choice(h * w, rand_sample, p=prob_vec, replace=False))         # util.py:247
    return indices
