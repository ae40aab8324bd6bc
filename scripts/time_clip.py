"""GPU: times the CLIP leg of bench.py (encode + cosine top-6 of a 4096 bank) per precision / batch; prints JSON lines."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from shapeclipper_b200 import clip  # noqa: E402

dev = torch.device("cuda:0")
pk = bench.peaks()
for B in [int(a) for a in sys.argv[1:]] or [16, 64]:
    for prec in clip.PRECISIONS:
        ctx = clip.bench_context(None, B, dev, precision=prec)
        ms = bench._time_cuda(lambda: ctx.run(ctx.images[0]), 20)
        ms_enc = bench._time_cuda(lambda: ctx.model.encode(ctx.images[0]), 20)
        tf = bench.CLIP_GFLOP_PER_IMAGE * B / ms_enc
        print(json.dumps(dict(batch=B, precision=prec, ms_encode_topk=ms, ms_encode=ms_enc, tflops=tf, frac=tf / pk["bf16_sustained"])), flush=True)
        del ctx
        torch.cuda.empty_cache()
