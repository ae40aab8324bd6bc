"""Marching-cubes case tables, GENERATED (not transcribed): for each of the 256 corner-sign configurations of a cell the triangles of
the iso-surface as triples of cell-edge ids. PyMCubes (the reference's `mcubes.marching_cubes`, utils/eval_3D.py:125) is a third-party
dependency that is absent here, so the tables are derived from the definition:

  * corner i of a cell sits at offset (i & 1, (i >> 1) & 1, (i >> 2) & 1); a corner is INSIDE when its value < isovalue;
  * an edge is crossed when its two corners differ; on every face the crossed edges are joined pairwise into iso-contour
    segments — an ambiguous face (4 crossed edges) joins the two edges around each INSIDE corner, a rule that depends only on the
    face's own corners, so the two cells sharing a face agree and the mesh is watertight;
  * the segments close into loops over the crossed edges; each loop of k edge points is triangulated as a fan (k - 2 triangles).

Edge e joins EDGE_CORNERS[e] = (lower corner, higher corner): interpolating from the lower corner gives both cells sharing an edge
the bit-identical vertex. Triangle orientation is consistent within a loop but not globally (the point-cloud metrics are
orientation-free)."""
import numpy as np

CORNER_OFFSETS = np.array([[i & 1, (i >> 1) & 1, (i >> 2) & 1] for i in range(8)], dtype=np.int32)
EDGE_CORNERS = np.array([(a, b) for a in range(8) for b in range(a + 1, 8) if bin(a ^ b).count("1") == 1], dtype=np.int32)   # 12 edges
_EDGE_ID = {(int(a), int(b)): e for e, (a, b) in enumerate(EDGE_CORNERS)}


def _edge(a, b):
    return _EDGE_ID[(min(a, b), max(a, b))]


def _faces():
    """6 faces, each as its 4 corners in cyclic order."""
    faces = []
    for axis in range(3):
        u, v = [ax for ax in range(3) if ax != axis]
        for side in (0, 1):
            cyc = []
            for du, dv in ((0, 0), (1, 0), (1, 1), (0, 1)):
                off = [0, 0, 0]
                off[axis], off[u], off[v] = side, du, dv
                cyc.append(off[0] | (off[1] << 1) | (off[2] << 2))
            faces.append(cyc)
    return faces


FACES = _faces()


def case_loops(inside):
    """inside: 8 booleans -> list of loops, each a list of edge ids in walking order."""
    adj = {}

    def link(e0, e1):
        adj.setdefault(e0, []).append(e1)
        adj.setdefault(e1, []).append(e0)
    for cyc in FACES:
        crossed = [(k, _edge(cyc[k], cyc[(k + 1) % 4])) for k in range(4) if inside[cyc[k]] != inside[cyc[(k + 1) % 4]]]
        if len(crossed) == 2:
            link(crossed[0][1], crossed[1][1])
        elif len(crossed) == 4:                              # ambiguous face: cut off each inside corner
            for k in range(4):
                if inside[cyc[k]]:
                    link(_edge(cyc[(k - 1) % 4], cyc[k]), _edge(cyc[k], cyc[(k + 1) % 4]))
    loops, seen = [], set()
    for start in sorted(adj):
        if start in seen:
            continue
        loop, prev, cur = [start], None, start
        seen.add(start)
        while True:
            nxt = [e for e in adj[cur] if e != prev]
            # degree is exactly 2; with both neighbours equal to prev (2-cycle) the walk ends
            step = nxt[0] if nxt else adj[cur][0]
            if len(adj[cur]) == 2 and adj[cur][0] == adj[cur][1]:
                step = adj[cur][0]
            if step == start or step in seen:
                break
            loop.append(step)
            seen.add(step)
            prev, cur = cur, step
        loops.append(loop)
    return loops


def build_tables():
    """-> (tri_count [256] int32, tri_edges [256, 3 * max_tris] int32 padded with -1, max_tris)."""
    tris = []
    for case in range(256):
        inside = [bool((case >> i) & 1) for i in range(8)]
        t = []
        for loop in case_loops(inside):
            for k in range(1, len(loop) - 1):
                t.append((loop[0], loop[k], loop[k + 1]))
        tris.append(t)
    max_tris = max(len(t) for t in tris)
    count = np.array([len(t) for t in tris], dtype=np.int32)
    edges = -np.ones((256, 3 * max_tris), dtype=np.int32)
    for case, t in enumerate(tris):
        flat = [e for tri in t for e in tri]
        edges[case, :len(flat)] = flat
    return count, edges, max_tris


TRI_COUNT, TRI_EDGES, MAX_TRIS = build_tables()
