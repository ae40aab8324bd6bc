import os, sys, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import options, render_fn, synthetic, _render_native as rn
from shapeclipper_b200.graph import HotPathGraph
dev = torch.device("cuda:0")
opt = options.default_options()
torch.manual_seed(0)
g = HotPathGraph(opt).to(dev)
batch = synthetic.make_batch(opt, 16, seed=1)
var, _ = synthetic.to_device(batch, dev)
render_fn.set_precision(forward="tc")
trace = torch.zeros(256, dtype=torch.int64, device=dev)
orig = rn.launch_forward
def patched(args, device, tc=False):
    if args.mode == 0:
        args.points_bar = ctypes.c_void_p(trace.data_ptr())
    return orig(args, device, tc=tc)
rn.launch_forward = patched
import shapeclipper_b200.render_fn as rf
for _ in range(2):
    trace.zero_()
    with torch.no_grad():
        g.renderer(opt, var.pose, var.intr, var.scale_dist, var.proj_latent_sdf, var.proj_latent_rgb, ray_idx=var.ray_idx, training=False)
    torch.cuda.synchronize()
t = trace.cpu()
for name, off in (("thread0", 0), ("thread300", 128)):
    v = t[off:off + 120]
    v = v[v > 0]
    d = (v[1:] - v[:-1]).tolist()
    print(name, len(v), "marks; deltas:", d)
