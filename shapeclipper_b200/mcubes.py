"""Iso-surface extraction and surface sampling on the GPU (csrc/mcubes.cu), with the call surface the reference uses from two
third-party packages that are absent here (utils/eval_3D.py:123-153):

    vertices, faces = mcubes.marching_cubes(level_vox_i, isovalue)          # PyMCubes: index-space vertices, triangle indices
    mesh = trimesh.Trimesh(vertices, faces); mesh.triangles; mesh.sample(opt.eval.num_points)

`shim.install()` registers this module as `mcubes` and `trimesh` when the real packages are not importable. The device-resident
entry points (`extract_triangles`, `sample_surface`) are what this package's own eval_3D.py uses: the level grid never leaves HBM.
Case tables: mcubes_tables.py (generated from the definition). Sampling: faces drawn with probability proportional to area by
inverting the cumulative area with searchsorted, a uniform point per face from two folded uniforms (trimesh.sample.sample_surface).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from . import mcubes_tables as tables

_vp = ctypes.c_void_p
_tables_on = set()


def declare(L):
    i, i64, f = ctypes.c_int, ctypes.c_int64, ctypes.c_float
    L.sc_mc_set_tables.argtypes = [_vp, _vp, _vp]
    L.sc_mc_set_tables.restype = i
    L.sc_mc_count.argtypes = [_vp, i, i, f, _vp, _vp]
    L.sc_mc_count.restype = i
    L.sc_mc_emit.argtypes = [_vp, i, i, f, _vp, f, f, _vp, _vp]
    L.sc_mc_emit.restype = i
    L.sc_tri_area.argtypes = [_vp, i64, _vp, _vp]
    L.sc_tri_area.restype = i
    L.sc_tri_sample.argtypes = [_vp, _vp, _vp, i64, _vp, _vp]
    L.sc_tri_sample.restype = i


def _ensure_tables(device):
    L = _lib.lib()
    key = torch.device(device).index or 0
    if key in _tables_on:
        return L
    if tables.MAX_TRIS > 5:
        raise RuntimeError("case table wider than the kernel's 5 triangles per cell")
    cnt = np.ascontiguousarray(tables.TRI_COUNT.astype(np.int8))
    edg = -np.ones((256, 15), dtype=np.int8)
    edg[:, :tables.TRI_EDGES.shape[1]] = tables.TRI_EDGES.astype(np.int8)
    ec = np.ascontiguousarray(tables.EDGE_CORNERS.astype(np.int8))
    with torch.cuda.device(device):
        _lib.check(L.sc_mc_set_tables(cnt.ctypes.data_as(_vp), np.ascontiguousarray(edg).ctypes.data_as(_vp), ec.ctypes.data_as(_vp)),
                   "sc_mc_set_tables")
    _tables_on.add(key)
    return L


@torch.no_grad()
def extract_triangles(level, isovalue=0.0, lo=0.0, hi=None):
    """level [B, n, n, n] fp32 CUDA -> list of B tensors [T_b, 3, 3]: the iso-surface triangles of each grid, vertices at
    index / n * (hi - lo) + lo (hi defaults to n: index units, as PyMCubes returns them)."""
    _lib.require_cuda(level)
    lv = level.float().contiguous()
    B, n = lv.shape[0], lv.shape[1]
    assert lv.dim() == 4 and lv.shape[2] == n and lv.shape[3] == n
    hi = float(n) if hi is None else float(hi)
    L = _ensure_tables(lv.device)
    cells = (n - 1) ** 3
    counts = torch.empty(B * cells, dtype=torch.int32, device=lv.device)
    from . import _render_native as rn
    with torch.cuda.device(lv.device):
        _lib.check(L.sc_mc_count(_lib.ptr(lv), B, n, float(isovalue), _lib.ptr(counts), _lib.stream_of(lv)), "sc_mc_count")
        incl = torch.cumsum(counts, 0, dtype=torch.int64)
        offsets = incl - counts
        per_grid = incl[cells - 1::cells].cpu()                        # the one host read: output sizes
        total = int(per_grid[-1])
        tris = torch.empty(max(total, 1), 3, 3, device=lv.device)
        if total:
            _lib.check(L.sc_mc_emit(_lib.ptr(lv), B, n, float(isovalue), _lib.ptr(offsets), float(lo), hi, _lib.ptr(tris),
                                    _lib.stream_of(lv)), "sc_mc_emit")
    rn.TIMERS.count(2)
    ends = per_grid.tolist()
    starts = [0] + ends[:-1]
    return [tris[s:e] for s, e in zip(starts, ends)]


@torch.no_grad()
def sample_surface(triangles, count, generator=None):
    """count points on the triangle soup [T, 3, 3] (CUDA), area-weighted; zeros [count, 3] for an empty mesh, as the reference
    does (utils/eval_3D.py:150-152). Draws: count face uniforms then count x 2 barycentric uniforms from `generator` (CUDA)."""
    dev = triangles.device
    T = triangles.shape[0]
    if T == 0:
        return torch.zeros(count, 3, device=dev)
    L = _lib.lib()
    tri = triangles.float().contiguous()
    area = torch.empty(T, device=dev)
    pts = torch.empty(count, 3, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.sc_tri_area(_lib.ptr(tri), T, _lib.ptr(area), _lib.stream_of(tri)), "sc_tri_area")
        cum = torch.cumsum(area.double(), 0)
        u = torch.rand(count, device=dev, generator=generator, dtype=torch.float64) * cum[-1]
        face = torch.searchsorted(cum, u).clamp_(max=T - 1)
        uv = torch.rand(count, 2, device=dev, generator=generator)
        _lib.check(L.sc_tri_sample(_lib.ptr(tri), _lib.ptr(face), _lib.ptr(uv), count, _lib.ptr(pts), _lib.stream_of(tri)), "sc_tri_sample")
    from . import _render_native as rn
    rn.TIMERS.count(2)
    return pts


# ---- the third-party call surface (numpy in / numpy out), for the reference's own eval_3D under shim.install()
def marching_cubes(volume, isovalue):
    """PyMCubes' signature: volume [n,n,n] array -> (vertices [V,3] float64 in index units, triangles [F,3] int64).
    Vertices are not shared between triangles (V = 3 F): everything downstream (trimesh.Trimesh(..).sample) is index-agnostic."""
    dev = torch.device("cuda", torch.cuda.current_device())
    lv = torch.as_tensor(np.ascontiguousarray(volume), dtype=torch.float32, device=dev)[None]
    tris = extract_triangles(lv, isovalue)[0]
    v = tris.reshape(-1, 3).double().cpu().numpy()
    return v, np.arange(v.shape[0], dtype=np.int64).reshape(-1, 3)


class Trimesh:
    """The slice of trimesh.Trimesh the reference touches: construction from (vertices, faces), `.triangles`, `.sample(count)`."""

    def __init__(self, vertices=None, faces=None, **_):
        self.vertices = np.zeros((0, 3)) if vertices is None else np.asarray(vertices, dtype=np.float64)
        self.faces = np.zeros((0, 3), dtype=np.int64) if faces is None else np.asarray(faces, dtype=np.int64)

    @property
    def triangles(self):
        return self.vertices[self.faces] if len(self.faces) else np.zeros((0, 3, 3))

    def sample(self, count):
        dev = torch.device("cuda", torch.cuda.current_device())
        tri = torch.as_tensor(self.triangles, dtype=torch.float32, device=dev)
        g = torch.Generator(device=dev)
        g.manual_seed(int(np.random.randint(0, 2 ** 31 - 1)))          # trimesh draws from numpy's global generator: stay seedable by it
        return sample_surface(tri, int(count), g).double().cpu().numpy()
