"""GPU: tensor-core render forward vs FP32 FFMA forward on a default-size batch + timing."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import options, render_fn, synthetic
from shapeclipper_b200.graph import HotPathGraph
dev = torch.device("cuda:0")
opt = options.default_options()
torch.manual_seed(0)
g = HotPathGraph(opt).to(dev)
with torch.no_grad():
    for p in g.sdf_network.parameters():
        p.add_(0.02 * torch.randn_like(p))
batch = synthetic.make_batch(opt, 16, seed=1)
var, _ = synthetic.to_device(batch, dev)
outs = {}
for mode in ("fp32", "tc"):
    render_fn.set_precision(forward=mode)
    torch.manual_seed(3)
    with torch.no_grad():
        o = g.renderer(opt, var.pose, var.intr, var.scale_dist, var.proj_latent_sdf, var.proj_latent_rgb, ray_idx=var.ray_idx, training=True)
    outs[mode] = o
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        with torch.no_grad():
            g.renderer(opt, var.pose, var.intr, var.scale_dist, var.proj_latent_sdf, var.proj_latent_rgb, ray_idx=var.ray_idx, training=False)
    e.record(); torch.cuda.synchronize()
    print(mode, "eval render ms", s.elapsed_time(e) / 5)
for nm, a, b in zip(["rgb", "mask", "mask_hard", "depth", "normal", "grad_eik"], outs["fp32"], outs["tc"]):
    print(nm, "max abs diff tc vs fp32: %.3e  (max |ref| %.3e) nan=%s" % (float((a - b).abs().max()), float(a.abs().max()), bool(torch.isnan(b).any())))
from shapeclipper_b200 import _render_native as rn
for mode in ("fp32", "tc"):
    render_fn.set_precision(forward=mode)
    rn.TIMERS.reset(); rn.TIMERS.enabled = True
    for _ in range(5):
        with torch.no_grad():
            g.renderer(opt, var.pose, var.intr, var.scale_dist, var.proj_latent_sdf, var.proj_latent_rgb, ray_idx=var.ray_idx, training=False)
    print(mode, {k: v[0] / v[1] for k, v in rn.TIMERS.totals_ms().items()})
    rn.TIMERS.enabled = False
