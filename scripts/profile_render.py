"""GPU: a few training renders (forward + backward) at bench.py's shape (batch 16, 512 rays x 64 samples) for ncu:

  ncu --set full --import-source on --clock-control none -k regex:render_tc_(fwd|bwd)_kernel --launch-skip 4 -c 2 \
      -o gpurun_out/render_tc python scripts/profile_render.py
"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import options, synthetic, _render_native as rn
from shapeclipper_b200.graph import HotPathGraph

B = int(os.environ.get("SC_PROFILE_BATCH", "16"))
dev = torch.device("cuda:0")
opt = options.default_options()
opt.loss_weight.nearest_img = None      # one render per step: launches alternate fwd<0>, (sdf query), bwd<0>
opt.loss_weight.nearest_mask = None
torch.manual_seed(0)
g = HotPathGraph(opt).to(dev)
with torch.no_grad():
    for p in g.sdf_network.parameters():
        p.add_(0.02 * torch.randn_like(p))
var, _ = synthetic.to_device(synthetic.make_batch(opt, B, seed=1), dev)
rn.TIMERS.enabled = True
for it in range(4):
    for p in g.parameters():
        p.grad = None
    _, loss = g(opt, var, training=True, get_loss=True)
    loss["all"].backward()
torch.cuda.synchronize()
print({k: (v[0] / v[1], v[1]) for k, v in rn.TIMERS.totals_ms().items()})
