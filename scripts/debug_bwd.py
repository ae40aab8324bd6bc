import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_render_gpu import _build, _load, NAMES
gd = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
for name, seed in (("render_train_full_8x8", 103), ("render_train_40rays", 102)):
    fx = _load(gd, name)
    opt, sdf, rgb, ren = _build(fx, fx["H"], fx["W"])
    leaves = {k: v.cuda().requires_grad_(True) for k, v in fx["inputs"].items()}
    ridx = fx["ray_idx"].cuda() if fx["ray_idx"] is not None else None
    for which in (["rgb"], ["mask"], ["depth"], ["normal"], ["grad_eik"], list(fx["cotangents"].keys())):
        torch.manual_seed(seed)
        out = ren(opt, leaves["pose"], leaves["intr"], leaves["scale_dist"], leaves["z_sdf"], leaves["z_rgb"], ray_idx=ridx, training=True)
        got = dict(zip(NAMES, out))
        scalar = sum((fx["cotangents"][n].cuda() * got[n]).sum() for n in which)
        g = torch.autograd.grad(scalar, [ren.density.beta], allow_unused=True)[0]
        print(name, which, "beta_bar", None if g is None else g.item(), "(ref all: %.5f)" % fx["grads"]["density.beta"].item())
