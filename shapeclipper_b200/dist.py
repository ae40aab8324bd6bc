"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch), ONE flat fp32
gradient buffer and ONE all-reduce per step in place of the reference's
DistributedDataParallel buckets (model/runner.py:113-121, utils/util.py:250-255)."""
import os

import torch
import torch.distributed as dist


def setup(backend=None):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun). No-op when single."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, 0
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def cleanup():
    if dist.is_initialized():
        dist.destroy_process_group()


class FlatGradients:
    """All parameters' gradients live as views into one flat fp32 buffer (plus `extra` floats standing for the
    parameters of the out-of-scope CNN encoders: 36 800 589 in total for the reference Graph), so the data-parallel
    exchange is a single sum all-reduce followed by a 1/world scale. Unused parameters keep a zero gradient, as under
    DDP(find_unused_parameters=True)."""

    def __init__(self, params, extra=0, device=None):
        self.params = [p for p in params]
        device = device or self.params[0].device
        n = sum(p.numel() for p in self.params) + int(extra)
        self.flat = torch.zeros(n, dtype=torch.float32, device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.extra = self.flat[off:]
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def zero(self):
        self.flat.zero_()

    def all_reduce(self):
        """Mean over ranks, in place: ONE collective. NCCL's AVG folds the 1/world scale into the reduction (no separate
        scaling kernel over the 147 MB buffer); gloo (CPU tests) has no AVG, so it sums and scales."""
        if self.world > 1:
            if dist.get_backend() == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
                self.flat.mul_(1.0 / self.world)

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            with torch.no_grad():
                for p in self.params:
                    dist.broadcast(p, src=src)      # in-place on the parameter itself (keeps version counters honest)
