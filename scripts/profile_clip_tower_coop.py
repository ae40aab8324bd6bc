"""GPU: cooperative CLIP tower encodes for an `ncu --set full --import-source on -k regex:clip_tower -s 3 -c 1` capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import clip  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
vis = clip.CLIPVisual("ViT-B/32", precision=prec).cuda()
img = torch.randn(B, 3, 224, 224, device="cuda")
for _ in range(5):
    vis.encode(img)
torch.cuda.synchronize()
