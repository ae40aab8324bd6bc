"""ORACLE (test infrastructure, not product code): CPU restatement of iso-surface extraction + surface sampling as the reference
uses them (utils/eval_3D.py:123-153: `mcubes.marching_cubes(level, isovalue)` -> vertices / S * (hi - lo) + lo ->
`trimesh.Trimesh(..).sample(n)`).

PARITY UNPINNED against PyMCubes / trimesh themselves: both are third-party dependencies of the reference (requirements.yaml), absent
from /root/reference and from this image, and the reference holds no golden meshes. What is pinned instead:
  * this file is an independent implementation of the same DEFINITION (no 256-entry table: every cell's iso-contour is traced on
    the cell's faces at run time), against which the GPU path must agree triangle for triangle;
  * closed-form facts the tests check on both: a sphere's mesh is watertight (every edge in exactly two triangles), its area tends to
    4 pi r^2, every vertex lies on a lattice edge at the linear zero crossing.

Definition (shared with shapeclipper_b200/mcubes_tables.py, stated there): a lattice point is inside when value < isovalue; crossed
cell edges get one vertex at the linear crossing, interpolated from the endpoint with the lower lattice index; on each cell face
crossed edges are joined pairwise, an ambiguous face joins the two edges around each inside corner; loops are fan-triangulated from
their lowest-numbered edge.
"""
import itertools

import numpy as np


def _cell_triangles(v, iso):
    """v[dx][dy][dz] (2x2x2 corner values) -> list of triangles, each 3 points in cell-local coordinates (float32 arithmetic)."""
    inside = {c: bool(v[c] < iso) for c in itertools.product((0, 1), repeat=3)}

    def key(a, b):                 # an edge = its two corners, lower lattice index first (z most significant, as in the tables)
        ka = a[0] + 2 * a[1] + 4 * a[2]
        kb = b[0] + 2 * b[1] + 4 * b[2]
        return (a, b) if ka < kb else (b, a)
    adj = {}

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)
    for axis in range(3):
        others = [ax for ax in range(3) if ax != axis]
        for side in (0, 1):
            cyc = []
            for du, dv in ((0, 0), (1, 0), (1, 1), (0, 1)):
                c = [0, 0, 0]
                c[axis], c[others[0]], c[others[1]] = side, du, dv
                cyc.append(tuple(c))
            crossed = [k for k in range(4) if inside[cyc[k]] != inside[cyc[(k + 1) % 4]]]
            if len(crossed) == 2:
                link(key(cyc[crossed[0]], cyc[(crossed[0] + 1) % 4]), key(cyc[crossed[1]], cyc[(crossed[1] + 1) % 4]))
            elif len(crossed) == 4:
                for k in range(4):
                    if inside[cyc[k]]:
                        link(key(cyc[(k - 1) % 4], cyc[k]), key(cyc[k], cyc[(k + 1) % 4]))

    def order(e):                  # edge numbering of the tables: pairs (a, b), a < b, lexicographic
        (a, b) = e
        return (a[0] + 2 * a[1] + 4 * a[2], b[0] + 2 * b[1] + 4 * b[2])

    def point(e):
        (a, b) = e
        va, vb = np.float32(v[a]), np.float32(v[b])
        t = (np.float32(iso) - va) / (vb - va)
        return [np.float32(a[i]) + t * np.float32(b[i] - a[i]) for i in range(3)]
    tris, seen = [], set()
    for start in sorted(adj, key=order):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nb = adj[cur]
            step = nb[0] if nb[0] != prev else nb[1]
            if nb[0] == nb[1]:
                step = nb[0]
            if step == start or step in seen:
                break
            loop.append(step)
            seen.add(step)
            prev, cur = cur, step
        pts = [point(e) for e in loop]
        for k in range(1, len(pts) - 1):
            tris.append([pts[0], pts[k], pts[k + 1]])
    return tris


def marching_cubes(level, iso=0.0, lo=0.0, hi=None):
    """level [n,n,n] -> triangles [T,3,3] float32 (cells in x-major, z-fastest order; vertices index / n * (hi - lo) + lo) and
    counts [(n-1)^3] int (triangles per cell, same order)."""
    lv = np.asarray(level, dtype=np.float32)
    n = lv.shape[0]
    hi = float(n) if hi is None else float(hi)
    scale = np.float32((hi - lo) / n)
    out, counts = [], []
    for x in range(n - 1):
        for y in range(n - 1):
            for z in range(n - 1):
                v = lv[x:x + 2, y:y + 2, z:z + 2]
                if (v < iso).all() or not (v < iso).any():
                    counts.append(0)
                    continue
                tris = _cell_triangles({c: v[c] for c in itertools.product((0, 1), repeat=3)}, iso)
                counts.append(len(tris))
                base = np.array([x, y, z], dtype=np.float32)
                for t in tris:
                    out.append([(np.array(p, dtype=np.float32) + base) * scale + np.float32(lo) for p in t])
    tri = np.array(out, dtype=np.float32).reshape(-1, 3, 3)
    return tri, np.array(counts, dtype=np.int64)


def triangle_areas(tri):
    u, w = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    return 0.5 * np.linalg.norm(np.cross(u, w), axis=1)


def sample_points(tri, face, uv):
    """trimesh.sample.sample_surface's point construction for given face indices and uniforms uv [count,2]."""
    uv = np.array(uv, dtype=np.float32).copy()
    fold = uv.sum(1) > 1.0
    uv[fold] = np.abs(uv[fold] - 1.0)
    a, b, c = tri[face, 0], tri[face, 1], tri[face, 2]
    return a + uv[:, :1] * (b - a) + uv[:, 1:] * (c - a)
