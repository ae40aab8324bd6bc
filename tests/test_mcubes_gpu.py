"""GPU: iso-surface extraction + surface sampling (csrc/mcubes.cu) against the CPU oracle (oracle/mcubes_ref.py; parity with
PyMCubes / trimesh themselves is unpinned, both are absent): triangle counts per cell index-exact, vertices to 1e-6, sampled points
from injected draws to 1e-6; closed-form checks at evaluate.py's full size (vox_res = 100); the reference's own eval_metrics
(utils/eval_3D.py:52-103) under shim.install()."""
import numpy as np
import pytest
import torch

from oracle import mcubes_ref as M

pytestmark = pytest.mark.gpu


def _field(n, seed):
    g = torch.linspace(-0.6, 0.6, n)
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    gen = torch.Generator().manual_seed(seed)
    c = (torch.rand(3, 3, generator=gen) - 0.5) * 0.5
    f = torch.full_like(X, 10.0)
    for i in range(3):                                             # union of three spheres: saddles and ambiguous faces occur
        f = torch.minimum(f, ((X - c[i, 0]) ** 2 + (Y - c[i, 1]) ** 2 + (Z - c[i, 2]) ** 2).sqrt() - (0.15 + 0.05 * i))
    return f + 0.01 * torch.randn(n, n, n, generator=gen)          # noise: every one of the 256 cases shows up


@pytest.mark.parametrize("n,seed,iso", [(9, 0, 0.0), (14, 1, 0.0), (12, 2, 0.02), (6, 3, -0.01)])
def test_triangles_match_the_oracle_exactly(n, seed, iso):
    from shapeclipper_b200 import mcubes
    lv = torch.stack([_field(n, seed), _field(n, seed + 10)])
    got = mcubes.extract_triangles(lv.cuda(), iso, lo=-0.6, hi=0.6)
    for b in range(2):
        want, counts = M.marching_cubes(lv[b].numpy(), iso, lo=-0.6, hi=0.6)
        g = got[b].cpu().numpy()
        assert g.shape == want.shape and want.shape[0] == counts.sum() > 0
        assert np.abs(g - want).max() <= 1e-6
    # per-cell counts through the C ABI
    L = mcubes._lib.lib()
    counts_dev = torch.empty(2 * (n - 1) ** 3, dtype=torch.int32, device="cuda")
    lvc = lv.cuda().contiguous()
    assert L.sc_mc_count(mcubes._lib.ptr(lvc), 2, n, float(iso), mcubes._lib.ptr(counts_dev), mcubes._lib.stream_of(lvc)) == 0
    want0 = M.marching_cubes(lv[0].numpy(), iso)[1]
    assert np.array_equal(counts_dev[:(n - 1) ** 3].cpu().numpy(), want0)


def test_sampling_matches_the_oracle_for_injected_draws_and_is_area_weighted():
    from shapeclipper_b200 import _lib, mcubes
    lv = _field(12, 5)
    tri = mcubes.extract_triangles(lv[None].cuda(), 0.0, lo=-0.6, hi=0.6)[0]
    T = tri.shape[0]
    g = torch.Generator().manual_seed(1)
    face = torch.randint(0, T, (4000,), generator=g)
    uv = torch.rand(4000, 2, generator=g)
    pts = torch.empty(4000, 3, device="cuda")
    L = _lib.lib()
    fc, uc = face.cuda(), uv.cuda()
    assert L.sc_tri_sample(_lib.ptr(tri), _lib.ptr(fc), _lib.ptr(uc), 4000, _lib.ptr(pts), _lib.stream_of(tri)) == 0
    want = M.sample_points(tri.cpu().numpy(), face.numpy(), uv.numpy())
    assert np.abs(pts.cpu().numpy() - want).max() <= 1e-6
    area = torch.empty(T, device="cuda")
    assert L.sc_tri_area(_lib.ptr(tri), T, _lib.ptr(area), _lib.stream_of(tri)) == 0
    assert np.allclose(area.cpu().numpy(), M.triangle_areas(tri.cpu().numpy()), rtol=1e-5, atol=1e-9)
    # area weighting: the share of samples on the upper half of the surface = its share of the area (binomial 5 sigma)
    gen = torch.Generator(device="cuda").manual_seed(3)
    p = mcubes.sample_surface(tri, 200000, gen)
    upper = tri.mean(1)[:, 2] > 0
    share = float(area[upper].sum() / area.sum())
    got = float((p[:, 2] > 0).float().mean())
    assert abs(got - share) < 5 * (share * (1 - share) / 200000) ** 0.5 + 2e-3
    assert mcubes.sample_surface(tri[:0], 7).abs().sum() == 0                          # empty mesh -> zeros (utils/eval_3D.py:150-152)


def test_full_size_sphere_properties():
    """vox_res = 100 (evaluate.py): 101^3 lattice, analytic sphere. Area -> 4 pi r^2, vertices on the sphere, samples uniform."""
    from shapeclipper_b200 import eval_3D, options
    opt = options.default_options(device="cuda:0")
    opt.eval.vox_res = 100
    n, r = 101, 0.4
    g = torch.linspace(-0.6, 0.6, n, device="cuda")
    X, Y, Z = torch.meshgrid(g, g, g, indexing="ij")
    level = ((X ** 2 + Y ** 2 + Z ** 2).sqrt() - r)[None].repeat(2, 1, 1, 1)
    meshes, clouds = eval_3D.convert_to_explicit(opt, level, 0.0, to_pointcloud=True, generator=torch.Generator(device="cuda").manual_seed(0))
    assert clouds.shape == (2, opt.eval.num_points, 3) and meshes[0].triangles.shape == meshes[1].triangles.shape
    s = n / (n - 1.0)                                                  # the reference's index / n scaling shrinks the lattice by (n-1)/n about lo
    tri = meshes[0].triangles
    centre = torch.tensor([-0.6 + 0.6 / s] * 3, device="cuda")
    rad = (tri.reshape(-1, 3) - centre).norm(dim=1)
    assert float((rad - r / s).abs().max()) < 2e-4
    from shapeclipper_b200 import _lib
    area = torch.empty(tri.shape[0], device="cuda")
    _lib.lib().sc_tri_area(_lib.ptr(tri.contiguous()), tri.shape[0], _lib.ptr(area), _lib.stream_of(tri))
    assert abs(float(area.sum()) / (4 * np.pi * (r / s) ** 2) - 1) < 2e-3
    d = (clouds[0] - centre)
    assert float((d.norm(dim=1) - r / s).abs().max()) < 2e-4
    assert float(d.mean(0).abs().max()) < 5e-3                          # uniform over the sphere: centroid at the centre
    octant = ((d > 0).long() * torch.tensor([1, 2, 4], device="cuda")).sum(1)
    frac = torch.bincount(octant, minlength=8).float() / d.shape[0]
    assert float((frac - 0.125).abs().max()) < 0.01


def test_reference_eval_metrics_runs_on_the_gpu_path_under_the_shim():
    """utils/eval_3D.eval_metrics (52-103) itself — mcubes / trimesh / chamfer_3D / the SDF network all provided by this package —
    against this package's eval_metrics on the same shapes: the geometry is identical, the random surface samples are not, so the
    chamfer / F-score agree to sampling noise."""
    import importlib
    import sys
    import refharness
    if not refharness.reference_available():
        pytest.skip("reference modules not staged")
    from shapeclipper_b200 import eval_3D as ours, options, shim
    for name in ("mcubes", "trimesh"):
        sys.modules.pop(name, None)
    shim.install()
    refharness.import_reference()
    ref_eval = importlib.reload(importlib.import_module("utils.eval_3D"))
    assert ref_eval.mcubes.__name__ == "shapeclipper_b200.mcubes" and ref_eval.trimesh.__name__ == "shapeclipper_b200.mcubes"
    from shapeclipper_b200.implicit import SDFNetwork
    opt = options.default_options(device="cuda:0")
    opt.eval.vox_res, opt.eval.num_points = 48, 20000
    torch.manual_seed(0)
    sdf = SDFNetwork(opt).cuda()
    B = 2
    g = torch.Generator().manual_seed(2)
    R = torch.eye(3).repeat(B, 1, 1)
    pose = torch.cat([R, torch.zeros(B, 3, 1)], -1).cuda()
    gt = torch.nn.functional.normalize(torch.randn(B, opt.eval.num_points, 3, generator=g), dim=-1).mul(0.5).cuda()
    edict = importlib.import_module("utils.util").EasyDict

    def make_var(cls):
        v = cls()
        v.idx = torch.arange(B, device="cuda")
        v.proj_latent_sdf = torch.zeros(B, 64, device="cuda")
        v.pose, v.pose_gt = pose.clone(), pose.clone()
        v.dpc = cls()
        v.dpc.points = gt.clone()
        return v
    va, vb = make_var(edict), make_var(options.Options)
    np.random.seed(0)
    acc_a, comp_a = ref_eval.eval_metrics(opt, va, sdf)
    acc_b, comp_b = ours.eval_metrics(opt, vb, sdf, generator=torch.Generator(device="cuda").manual_seed(0))
    assert va.dpc_pred.shape == vb.dpc_pred.shape == (B, opt.eval.num_points, 3)
    assert len(va.mesh_pred) == B and va.mesh_pred[0].triangles.shape[0] == vb.mesh_pred[0].triangles.shape[0] > 1000
    assert abs(float(acc_a) - float(acc_b)) < 0.02 * float(acc_b) + 1e-4 and abs(float(comp_a) - float(comp_b)) < 0.02 * float(comp_b) + 1e-4
    assert float((va.f_score - vb.f_score).abs().max()) < 0.02
