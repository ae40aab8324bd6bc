"""GPU: the boundary-distance ray sampler on one batch of 32 masks of 224 x 224 (bench.py's ray_sampler sub-record), for ncu:

  ncu --set full --clock-control none -k regex:edt_ -c 2 --launch-skip 4 -o gpurun_out/sampler python scripts/profile_sampler.py
"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import sampling
g = np.random.RandomState(0)
yy, xx = np.mgrid[0:224, 0:224]
m = np.stack([((yy - g.randint(60, 160)) ** 2 / float(g.randint(30, 80)) ** 2 + (xx - g.randint(60, 160)) ** 2 / float(g.randint(30, 80)) ** 2 < 1)
              for _ in range(32)]).astype(np.float32)
md = torch.from_numpy(m).cuda()
for _ in range(4):
    d = sampling.boundary_distance(md)
torch.cuda.synchronize()
print(float(d.max()))
