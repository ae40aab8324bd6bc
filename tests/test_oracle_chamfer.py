"""CPU: the C restatement of the chamfer kernel against closed-form / brute-force answers."""
import numpy as np

from oracle import chamfer_ref


def test_self_distance_is_zero_and_identity_index():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((2, 700, 3)).astype(np.float32)
    d1, d2, i1, i2 = chamfer_ref.chamfer_forward(a, a)
    assert (d1 == 0).all() and (d2 == 0).all()
    assert (i1 == np.arange(700)[None]).all() and (i2 == np.arange(700)[None]).all()


def test_lowest_index_wins_ties_across_tiles():
    rng = np.random.default_rng(1)
    b = rng.standard_normal((1, 1500, 3)).astype(np.float32)
    b[0, 1100] = b[0, 17]          # duplicate in a later 512-tile
    b[0, 600] = b[0, 17]
    a = b[:, [17]] + np.float32(1e-3)
    _, _, i1, _ = chamfer_ref.chamfer_forward(a, b)
    assert i1[0, 0] == 17


def test_matches_float64_bruteforce_on_exact_inputs():
    rng = np.random.default_rng(2)
    # small integers / 8: differences and squares are exact in fp32, so fp64 brute force is an exact check
    a = (rng.integers(-40, 40, (3, 257, 3)) / 8).astype(np.float32)
    b = (rng.integers(-40, 40, (3, 1031, 3)) / 8).astype(np.float32)
    d1, d2, i1, i2 = chamfer_ref.chamfer_forward(a, b)
    e1, j1 = chamfer_ref.chamfer_forward_numpy(a, b)
    e2, j2 = chamfer_ref.chamfer_forward_numpy(b, a)
    assert (d1 == e1).all() and (i1 == j1).all()
    assert (d2 == e2).all() and (i2 == j2).all()


def _round_f32(fr):
    """Exact rational -> nearest float32 (ties to even not needed for the values used here)."""
    from fractions import Fraction
    c = np.float32(float(fr))
    best = c
    for o in (np.nextafter(c, np.float32(np.inf)), np.nextafter(c, np.float32(-np.inf))):
        if abs(Fraction(float(o)) - fr) < abs(Fraction(float(best)) - fr):
            best = o
    return best


def _fma32(a, b, c):
    from fractions import Fraction
    return _round_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def test_fma_contraction_is_the_documented_one():
    # pairs where fmaf(z,z,fmaf(x,x,y*y)) differs from other contractions / the separately rounded sum
    rng = np.random.default_rng(4)
    n_diff = 0
    for _ in range(200):
        x, y, z = (rng.standard_normal(3) * 2).astype(np.float32)
        a = np.zeros((1, 1, 3), np.float32)
        b = np.array([[[x, y, z]]], np.float32)
        d1, _, _, _ = chamfer_ref.chamfer_forward(a, b)
        want = _fma32(z, z, _fma32(x, x, np.float32(y * y)))
        assert d1[0, 0] == want
        plain = np.float32(np.float32(np.float32(x * x) + np.float32(y * y)) + np.float32(z * z))
        n_diff += int(plain != want)
    assert n_diff > 0      # the test inputs do distinguish the two definitions


def test_backward_matches_analytic():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((2, 50, 3)).astype(np.float32)
    b = rng.standard_normal((2, 70, 3)).astype(np.float32)
    d1, d2, i1, i2 = chamfer_ref.chamfer_forward(a, b)
    g1 = rng.standard_normal(d1.shape).astype(np.float32)
    g2 = rng.standard_normal(d2.shape).astype(np.float32)
    ga, gb = chamfer_ref.chamfer_backward(a, b, g1, g2, i1, i2)
    ea, eb = np.zeros_like(a, np.float64), np.zeros_like(b, np.float64)
    for bb in range(2):
        for j in range(50):
            v = 2 * g1[bb, j] * (a[bb, j].astype(np.float64) - b[bb, i1[bb, j]])
            ea[bb, j] += v; eb[bb, i1[bb, j]] -= v
        for j in range(70):
            v = 2 * g2[bb, j] * (b[bb, j].astype(np.float64) - a[bb, i2[bb, j]])
            eb[bb, j] += v; ea[bb, i2[bb, j]] -= v
    assert np.allclose(ga, ea, atol=1e-5) and np.allclose(gb, eb, atol=1e-5)
