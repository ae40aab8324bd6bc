"""Generates tests/golden/ray_sampler.npz from the UNMODIFIED reference (build container only):

    python tests/gen_golden_sampler.py

`utils.util.compute_sampling_prob` (utils/util.py:237-248) is run with `vigra.filters.boundaryDistanceTransform` stubbed by the
scipy-based oracle transform (vigra is absent), under a fixed `np.random.seed`, on three masks: a soft 224 x 224 two-blob mask, a
48 x 80 noise mask (the function only asserts opt.H == h) and a single foreground pixel."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def masks():
    g = np.random.RandomState(11)
    out = {}
    yy, xx = np.mgrid[0:224, 0:224]
    m = ((yy - 90) ** 2 / 55.0 ** 2 + (xx - 100) ** 2 / 38.0 ** 2 < 1) | ((yy - 160) ** 2 + (xx - 170) ** 2 < 24 ** 2)
    out["blobs224"] = (m.astype(np.float32) * 0.9 + 0.05 * ((xx % 7) / 7.0).astype(np.float32))      # soft mask: 0..0.05 / 0.9..0.95
    r = g.rand(48, 80).astype(np.float32)
    out["noise48x80"] = (r > 0.7).astype(np.float32)
    one = np.zeros((32, 32), dtype=np.float32); one[5, 27] = 1.0
    out["onepixel32"] = one
    return out


def main():
    import refharness
    from oracle import sampling_ref
    refharness.import_reference()
    import vigra                      # the empty stub refharness registered
    import types
    vigra.filters = types.SimpleNamespace(boundaryDistanceTransform=lambda a: sampling_ref.boundary_distance_scipy(np.asarray(a) > 0.5))
    from utils import util
    util.vigra = vigra
    from shapeclipper_b200.options import Options as edict
    out = {}
    for name, m in masks().items():
        H, W = m.shape
        n = min(512, H * W // 4)
        opt = edict(H=H, W=W, render=edict(rand_sample=n))
        np.random.seed(1234)
        idx = util.compute_sampling_prob(opt, torch.from_numpy(m), 3)
        out[name + "_mask"] = m
        out[name + "_idx"] = idx.numpy()
        out[name + "_dist"] = sampling_ref.boundary_distance_scipy(m > 0.5)
    np.savez_compressed(os.path.join(HERE, "golden", "ray_sampler.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
