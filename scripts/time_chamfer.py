"""GPU: times sc_chamfer_forward against the recompiled reference kernel (oracle/_ref) at eval size."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapeclipper_b200 import chamfer_3D  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    out = []
    for (B, N, M) in [(1, 100000, 100000), (16, 100000, 100000), (1, 10000, 10000)]:
        g = torch.Generator().manual_seed(0)
        a = torch.randn(B, N, 3, generator=g).cuda()
        b = torch.randn(B, M, 3, generator=g).cuda()
        outs = [torch.zeros(B, N, device="cuda"), torch.zeros(B, M, device="cuda"),
                torch.zeros(B, N, dtype=torch.int32, device="cuda"), torch.zeros(B, M, dtype=torch.int32, device="cuda")]
        ms = timeit(lambda: chamfer_3D.forward(a, b, *outs))
        rec = dict(B=B, N=N, M=M, ours_ms=ms, ours_Tpairs_s=2.0 * B * N * M / ms / 1e9)
        try:
            from oracle import build_ref
            ref = build_ref.load()
            if ref is not None:
                ms_r = timeit(lambda: ref.forward(a, b, *outs), iters=3, warm=1)
                rec.update(ref_kernel_ms=ms_r, speedup_vs_ref_kernel=ms_r / ms)
        except Exception as ex:  # noqa: BLE001
            rec["ref_error"] = repr(ex)
        out.append(rec)
        print(json.dumps(rec), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/time_chamfer.json", "w"), indent=1)


if __name__ == "__main__":
    main()
