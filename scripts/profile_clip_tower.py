"""GPU: one CLIP tower encode as one launch per phase (same device code as the cooperative launch), for
  ncu --kernel-name regex:clip_tower --launch-skip N --launch-count N --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active ...
    python scripts/profile_clip_tower.py [batch] [split|fp16]      (prints N = phases per encode)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import clip  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
vis = clip.CLIPVisual("ViT-B/32", precision=prec).cuda()
img = torch.randn(B, 3, 224, 224, device="cuda")
vis.per_phase_launches = True
vis.encode(img)          # warm-up: N launches
torch.cuda.synchronize()
vis.encode(img)          # profiled: N launches
torch.cuda.synchronize()
print("phases per encode:", vis._tower[2][B].n_phases)
