"""CPU: the marching-cubes oracle (oracle/mcubes_ref.py) and the generated case tables (shapeclipper_b200/mcubes_tables.py) against
each other and against closed-form facts — PyMCubes / trimesh are absent, parity with them is unpinned (see the oracle's header)."""
import itertools

import numpy as np

from oracle import mcubes_ref as M
from shapeclipper_b200 import mcubes_tables as T


def _sphere(n, r=0.35, c=(0.03, -0.02, 0.01)):
    g = np.linspace(-0.6, 0.6, n, dtype=np.float32)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    return np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2) - r


def test_tables_agree_with_the_per_cell_oracle_on_all_256_cases():
    for case in range(256):
        v = {c: (-1.0 if (case >> (c[0] + 2 * c[1] + 4 * c[2])) & 1 else 1.0) for c in itertools.product((0, 1), repeat=3)}
        want = M._cell_triangles(v, 0.0)
        assert len(want) == T.TRI_COUNT[case]
        got = []
        for k in range(T.TRI_COUNT[case]):
            tri = []
            for e in T.TRI_EDGES[case, 3 * k:3 * k + 3]:
                a, b = T.CORNER_OFFSETS[T.EDGE_CORNERS[e][0]], T.CORNER_OFFSETS[T.EDGE_CORNERS[e][1]]
                tri.append(((a + b) / 2.0).tolist())                     # +-1 corner values: crossings at edge midpoints
            got.append(tri)
        assert np.allclose(np.array(got, dtype=np.float32).reshape(-1, 3, 3), np.array(want, dtype=np.float32).reshape(-1, 3, 3)), case


def test_complementary_cases_have_the_same_crossed_edges():
    for case in range(256):
        a = set(T.TRI_EDGES[case][T.TRI_EDGES[case] >= 0].tolist())
        b = set(T.TRI_EDGES[255 - case][T.TRI_EDGES[255 - case] >= 0].tolist())
        assert a == b


def test_sphere_mesh_is_watertight_and_has_the_right_area():
    n = 21
    tri, counts = M.marching_cubes(_sphere(n), 0.0, lo=-0.6, hi=0.6 * n / (n - 1) * 2 - 0.6)      # hi chosen so that index * scale + lo hits the lattice
    assert tri.shape[0] == counts.sum() > 500
    edges = {}
    for t in tri:
        k = [tuple(np.round(p, 5)) for p in t]
        for i in range(3):
            e = tuple(sorted((k[i], k[(i + 1) % 3])))
            edges[e] = edges.get(e, 0) + 1
    assert set(edges.values()) == {2}                                     # closed 2-manifold: every edge in exactly two triangles
    area = M.triangle_areas(tri).sum()
    assert abs(area - 4 * np.pi * 0.35 ** 2) / (4 * np.pi * 0.35 ** 2) < 0.03
    r = np.linalg.norm(tri.reshape(-1, 3) - np.array([0.03, -0.02, 0.01], dtype=np.float32), axis=1)
    assert np.abs(r - 0.35).max() < 0.01                                  # vertices sit on the sphere (linear interpolation of a distance field)


def test_sampler_points_lie_in_their_triangles():
    tri, _ = M.marching_cubes(_sphere(9), 0.0)
    rng = np.random.RandomState(0)
    face = rng.randint(0, tri.shape[0], size=200)
    uv = rng.rand(200, 2).astype(np.float32)
    p = M.sample_points(tri, face, uv)
    a, b, c = tri[face, 0], tri[face, 1], tri[face, 2]
    nrm = np.cross(b - a, c - a)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True) + 1e-20
    assert np.abs(((p - a) * nrm).sum(1)).max() < 1e-5                    # in the triangle's plane
    # barycentric coordinates within [0, 1]
    m = np.stack([b - a, c - a], -1)
    sol = np.stack([np.linalg.lstsq(m[i], (p - a)[i], rcond=None)[0] for i in range(200)])
    assert sol.min() > -1e-4 and sol.sum(1).max() < 1 + 1e-4
