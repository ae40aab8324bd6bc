"""GPU: the boundary-distance ray sampler (csrc/sampler.cu, shapeclipper_b200/sampling.py; utils/util.py:237-248) through the C ABI:
bit-equal distances against the oracle (scipy's exact EDT / the definition), the reference's own indices (golden, and the
reference's function itself under the shim), and the law of the batched device draw."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
GOLDEN = os.path.join(HERE, "golden", "ray_sampler.npz")
NAMES = ("blobs224", "noise48x80", "onepixel32")


def _dev():
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", NAMES)
def test_distance_bit_equal_to_golden(name):
    from shapeclipper_b200 import sampling
    z = np.load(GOLDEN)
    d = sampling.boundary_distance(torch.from_numpy(z[name + "_mask"]).to(_dev()))
    assert d.dtype == torch.float32 and np.array_equal(d.cpu().numpy(), z[name + "_dist"])


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 1, 40), (2, 33, 1), (4, 31, 65), (2, 64, 64), (32, 224, 224), (1, 300, 517)])
def test_distance_batched_ragged_sizes_vs_oracle(shape):
    from oracle import sampling_ref
    from shapeclipper_b200 import sampling
    g = np.random.RandomState(sum(shape))
    B, H, W = shape
    m = np.zeros(shape, dtype=np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):                                   # blobs of random size plus salt noise; image 0 all background, image 1 all foreground
        cy, cx, r = g.randint(0, H), g.randint(0, W), g.randint(1, max(2, min(H, W) // 2 + 1))
        m[b] = (((yy - cy) ** 2 + (xx - cx) ** 2 < r * r) ^ (g.rand(H, W) > 0.98)).astype(np.float32)
    if B > 1:
        m[0] = 0.0
        m[1] = 1.0
    d = sampling.boundary_distance(torch.from_numpy(m).to(_dev())).cpu().numpy()
    for b in range(B if B <= 4 else 6):
        want = sampling_ref.boundary_distance_scipy(m[b] > 0.5)
        assert np.array_equal(d[b], want), (b, np.abs(d[b] - want).max())
    if H * W <= 64 * 64 // 8:
        assert np.array_equal(d[-1], sampling_ref.boundary_distance_bruteforce(m[-1] > 0.5))


def test_threshold_and_non_binary_masks():
    from oracle import sampling_ref
    from shapeclipper_b200 import sampling
    g = np.random.RandomState(2)
    m = g.rand(50, 70).astype(np.float32)
    m[10, 10] = 0.5                                      # exactly 0.5 is background (mask > 0.5, util.py:241)
    d = sampling.boundary_distance(torch.from_numpy(m).to(_dev())).cpu().numpy()
    assert np.array_equal(d, sampling_ref.boundary_distance_scipy(m > 0.5))
    with pytest.raises(ValueError):
        sampling.boundary_distance(torch.from_numpy(m))  # no CPU path


@pytest.mark.parametrize("name", NAMES)
def test_compute_sampling_prob_reproduces_the_reference_indices(name):
    """The reference's function, its numpy draw, the GPU transform: the indices the reference itself produced."""
    from shapeclipper_b200 import sampling
    from shapeclipper_b200.options import Options
    z = np.load(GOLDEN)
    m, want = z[name + "_mask"], z[name + "_idx"]
    opt = Options(H=m.shape[0], W=m.shape[1], render=dict(rand_sample=len(want)))
    np.random.seed(1234)
    got = sampling.compute_sampling_prob(opt, torch.from_numpy(m), 3)
    assert got.dtype == torch.int64 and not got.is_cuda and np.array_equal(got.numpy(), want)
    np.random.seed(1234)
    assert np.array_equal(sampling.compute_sampling_prob(opt, torch.from_numpy(m).to(_dev()), 3).numpy(), want)


def test_reference_function_runs_on_the_shim():
    """utils.util.compute_sampling_prob itself (staged bytecode on the GPU box) with `vigra` resolved by shim.install()."""
    import refharness
    if not refharness.reference_available():
        pytest.skip("reference (or its staged bytecode) not present")
    refharness.import_reference()
    import shapeclipper_b200.shim as shim
    from shapeclipper_b200 import sampling
    from shapeclipper_b200.options import Options
    shim.install()
    assert sys.modules["vigra"] is sampling.vigra
    from utils import util
    old, util.vigra = util.vigra, sys.modules["vigra"]   # util imported `vigra` (the empty test stub) before the shim was installed
    try:
        z = np.load(GOLDEN)
        for name in NAMES:
            m, want = z[name + "_mask"], z[name + "_idx"]
            opt = Options(H=m.shape[0], W=m.shape[1], render=dict(rand_sample=len(want)))
            np.random.seed(1234)
            assert np.array_equal(util.compute_sampling_prob(opt, torch.from_numpy(m), 3).numpy(), want)
    finally:
        util.vigra = old


def test_sample_rays_no_duplicates_in_range_and_deterministic():
    from shapeclipper_b200 import sampling
    z = np.load(GOLDEN)
    m = torch.from_numpy(z["blobs224_mask"]).to(_dev())[None].repeat(5, 1, 1)
    gen = torch.Generator(device=_dev()).manual_seed(9)
    idx = sampling.sample_rays(m, 512, 3.0, generator=gen)
    assert idx.shape == (5, 512) and idx.dtype == torch.int64 and idx.is_cuda
    assert int(idx.min()) >= 0 and int(idx.max()) < 224 * 224
    for b in range(5):
        assert len(torch.unique(idx[b])) == 512
    assert not torch.equal(idx[0], idx[1])               # independent draws per image
    gen.manual_seed(9)
    assert torch.equal(idx, sampling.sample_rays(m, 512, 3.0, generator=gen))
    with pytest.raises(ValueError):
        sampling.sample_rays(m[:, :4, :4], 17)


def test_sample_rays_follows_the_reference_law():
    """First draw of each image ~ p = normalize(1 / (d + fac)); whole draws of n without replacement: inclusion frequencies against
    np.random.choice(replace=False) with the same p (the reference's draw), both from many repetitions."""
    from oracle import sampling_ref
    from shapeclipper_b200 import sampling
    g = np.random.RandomState(4)
    lab = np.zeros((6, 8), bool); lab[2:4, 3:6] = True
    p = 1.0 / (sampling_ref.boundary_distance_scipy(lab).astype(np.float64) + 1.0)
    p = (p / p.sum()).ravel()
    N = 40000
    m = torch.from_numpy(lab.astype(np.float32)).to(_dev())[None].repeat(N, 1, 1)
    idx = sampling.sample_rays(m, 6, 1.0, generator=torch.Generator(device=_dev()).manual_seed(1)).cpu().numpy()
    first = np.bincount(idx[:, 0], minlength=48) / N
    assert np.abs(first - p).max() < 5 * np.sqrt(p.max() / N)
    incl = np.bincount(idx.ravel(), minlength=48) / N
    ref = np.zeros(48)
    for _ in range(4000):
        ref[g.choice(48, 6, p=p, replace=False)] += 1
    ref /= 4000
    assert np.abs(incl - ref).max() < 5 * np.sqrt(0.25 / 4000)
