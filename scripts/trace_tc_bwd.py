"""GPU, library built with SC_TC_TRACE=1: clock64 marks of thread 0 (the MMA issuer) and thread 300 through the first tile(s)
of CTA 0 of render_tc_bwd_kernel<0>. Marks per layer GEMM: gemm() entry, after acquire (CTA barrier + weights landed);
per accumulator read: wait start, MMA done, TMEM load done."""
import os, sys, torch, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import options, synthetic, _render_native as rn
from shapeclipper_b200.graph import HotPathGraph
dev = torch.device("cuda:0")
opt = options.default_options()
opt.loss_weight.nearest_img = None
opt.loss_weight.nearest_mask = None
torch.manual_seed(0)
g = HotPathGraph(opt).to(dev)
var, _ = synthetic.to_device(synthetic.make_batch(opt, 16, seed=1), dev)
trace = torch.zeros(1024, dtype=torch.int64, device=dev)
orig = rn.launch_backward
def patched(args, device, tc=False):
    if args.mode == 0:
        trace.zero_()
        args.points_bar = ctypes.c_void_p(trace.data_ptr())
    return orig(args, device, tc=tc)
rn.launch_backward = patched
for _ in range(3):
    for p in g.parameters():
        p.grad = None
    _, loss = g(opt, var, training=True, get_loss=True)
    loss["all"].backward()
    torch.cuda.synchronize()
t = trace.cpu()
for name, off in (("thread0", 0), ("thread300", 512)):
    v = t[off:off + 500]
    v = v[v > 0]
    d = (v[1:] - v[:-1]).tolist()
    print(name, len(v), "marks; total", int(v[-1] - v[0]), "deltas:", d)
