"""CPU: the ray-sampler oracle (oracle/sampling_ref.py) against the definition, against scipy, against the frozen output of the
reference's own utils.util.compute_sampling_prob (tests/golden/ray_sampler.npz) and — where the reference is present — against
that function run live with vigra stubbed by the oracle transform."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from oracle import sampling_ref  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "ray_sampler.npz")
NAMES = ("blobs224", "noise48x80", "onepixel32")


def _cases():
    g = np.random.RandomState(3)
    yield "noise", g.rand(13, 17) > 0.6
    yield "stripes", (np.arange(12)[:, None] // 3 % 2 == 0) & np.ones((12, 9), bool)
    one = np.zeros((9, 11), bool); one[4, 0] = True
    yield "one_pixel_on_the_border", one
    hole = np.ones((10, 10), bool); hole[3:5, 6] = False
    yield "hole", hole
    yield "all_background", np.zeros((5, 7), bool)
    yield "all_foreground", np.ones((6, 4), bool)
    yield "single_pixel_image", np.ones((1, 1), bool)


@pytest.mark.parametrize("name,lab", list(_cases()), ids=[c[0] for c in _cases()])
def test_scipy_transform_equals_the_definition(name, lab):
    a = sampling_ref.boundary_distance_bruteforce(lab)
    b = sampling_ref.boundary_distance_scipy(lab)
    assert a.dtype == np.float32 and b.dtype == np.float32
    assert np.array_equal(a, b)
    if lab.any() and not lab.all():
        assert a.min() == np.float32(0.5)                 # a pixel next to the boundary: 1 - 0.5 (InterpixelBoundary)


def test_golden_distances_are_the_oracle():
    z = np.load(GOLDEN)
    for n in NAMES:
        assert np.array_equal(sampling_ref.boundary_distance_scipy(z[n + "_mask"] > 0.5), z[n + "_dist"])


@pytest.mark.parametrize("name", NAMES)
def test_restatement_reproduces_the_reference_indices(name):
    z = np.load(GOLDEN)
    m, want = z[name + "_mask"], z[name + "_idx"]
    np.random.seed(1234)
    got = sampling_ref.compute_sampling_prob(m.shape[0], len(want), torch.from_numpy(m), 3)
    assert got.dtype == torch.int64 and np.array_equal(got.numpy(), want)
    assert len(np.unique(want)) == len(want) and want.min() >= 0 and want.max() < m.size


def test_restatement_matches_live_reference():
    import refharness
    if not refharness.reference_available():
        pytest.skip("reference tree not present")
    refharness.import_reference()
    import types
    from utils import util
    from shapeclipper_b200.options import Options
    stub = types.SimpleNamespace(filters=types.SimpleNamespace(
        boundaryDistanceTransform=lambda a: sampling_ref.boundary_distance_scipy(np.asarray(a) > 0.5)))
    old, util.vigra = util.vigra, stub
    try:
        g = np.random.RandomState(5)
        m = torch.from_numpy((g.rand(40, 56) > 0.55).astype(np.float32) * 0.8 + 0.1)
        opt = Options(H=40, W=56, render=dict(rand_sample=200))
        np.random.seed(77)
        want = util.compute_sampling_prob(opt, m, 3)
        np.random.seed(77)
        got = sampling_ref.compute_sampling_prob(40, 200, m, 3)
        assert torch.equal(want, got)
        np.random.seed(78)
        assert torch.equal(util.compute_sampling_prob(opt, m, uniform_fac=1), (np.random.seed(78), sampling_ref.compute_sampling_prob(40, 200, m, 1))[1])
    finally:
        util.vigra = old
