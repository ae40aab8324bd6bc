"""GPU: a few CLIP ViT-B/32 encodes at bench.py's batch (16) for ncu / timing:
  ncu --set full -k regex:"gemm_tc_kernel|attention_kernel|layernorm_kernel" --launch-skip 100 -c 9 -o gpurun_out/clip python scripts/profile_clip.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import clip
B = int(os.environ.get("SC_PROFILE_BATCH", "16"))
m = clip.CLIPVisual("ViT-B/32", precision="split").cuda()
x = torch.randn(B, 3, 224, 224, device="cuda")
for _ in range(3):
    m.encode(x)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    m.encode(x)
e.record(); torch.cuda.synchronize()
print("clip encode ms", s.elapsed_time(e) / 10, "batch", B)
