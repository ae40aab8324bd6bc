"""One training iteration of the hot path, shaped like Runner.train_iteration (model/runner.py:235-292):

    optim.zero_grad() -> graph.forward(opt, var, training=True, get_loss=True) -> loss.all.backward()
    -> [flat gradient all-reduce, N > 1] -> optim.step()

The reference runs this as ~2 000 eager kernel launches with ~20 host syncs. Here the step is launch-bound once the
renders are fused (two render kernels each way + ~350 small torch kernels for rays, losses and Adam), so `TrainStep`
captures it ONCE into a CUDA graph — forward + losses + backward, the NCCL gradient all-reduce (averaging, N > 1) and the
optimiser step in ONE graph, so a step is a single replay with no host gap around the collective — and replays it: same
kernels, same arithmetic, no per-launch host cost. (`SC_ALLREDUCE_IN_GRAPH=0` keeps the collective eager between two
graphs, the round-1 structure.)

Requirements for capture (checked): `opt.render.device_rng` and `opt.reg.device_sampling` (every random draw on the
CUDA generator, no host round trip) and a capturable optimiser (`torch.optim.Adam(..., capturable=True)`).
Without them the step runs eagerly with identical results to calling the pieces by hand.
"""
import os

import torch

from . import _render_native as rn
from . import render_fn
from .options import Options

GRAD_LEAVES = ("pose", "intr", "scale_dist", "proj_latent_sdf", "proj_latent_rgb",
               "pose_NN", "intr_NN", "scale_dist_NN", "proj_latent_rgb_NN")


class TrainStep:
    def __init__(self, opt, graph, optim, flat_grads, example_batch, device, side_work=None, use_cuda_graph=True,
                 warmup=3, side_stream=True):
        """graph: HotPathGraph; flat_grads: dist.FlatGradients over the optimiser's parameters; example_batch: a host
        batch (synthetic.make_batch layout) fixing every shape; side_work: optional callable run at the start of each
        step on the same stream (bench.py: the CLIP encode + k-NN leg)."""
        self.opt, self.graph, self.optim, self.flat = opt, graph, optim, flat_grads
        self.device = torch.device(device)
        self.side_work = side_work
        # side_work (independent of the render path) runs on its own low-priority stream, forked at the start of the step and
        # joined after the backward: its kernels fill the SMs the small glue kernels between the render launches leave idle
        self._side = torch.cuda.Stream(self.device) if (side_work is not None and side_stream) else None
        self.var = Options()
        for k, t in example_batch.items():
            d = t.to(self.device)
            if k in GRAD_LEAVES:
                d.requires_grad_(True)
            self.var[k] = d
        self.loss = None
        self.launches_per_step = None
        self._g_main = self._g_optim = None
        self.use_cuda_graph = bool(use_cuda_graph)
        if self.use_cuda_graph:
            if not (getattr(opt.render, "device_rng", False) and getattr(opt.reg, "device_sampling", False)):
                raise ValueError("CUDA-graph capture needs opt.render.device_rng and opt.reg.device_sampling")
            if not all(g.get("capturable", False) for g in optim.param_groups):
                raise ValueError("CUDA-graph capture needs a capturable optimiser (Adam(..., capturable=True))")
            self._capture(warmup)

    # ------------------------------------------------------------------------------------------------------
    def load(self, batch):
        """Host (pinned) batch -> the step's static device tensors; async on the current stream."""
        with torch.no_grad():
            for k, t in batch.items():
                self.var[k].copy_(t, non_blocking=True)

    def run_epoch(self, batches, n_steps=None, extra=None):
        """Runner.train_epoch's loop (model/runner.py:198-225: `batch = next(loader); var = move_to_device(batch); loss =
        train_iteration(...)`, the progress bar reading `loss.all` on the host every iteration) over pinned host batches,
        software-pipelined the way a prefetching loader would be: batch i+1 travels host -> device (a staging copy, on a copy
        stream) while step i runs, and the loss of step i-1 is read on the host while step i runs. Every step's inputs are copied
        and every step's loss is read; nothing is skipped, only overlapped.
        batches: sequence of host batches (cycled); extra: optional callable i -> [(device_tensor, pinned_host_tensor), ...] of
        further per-step inputs (bench.py: the CLIP leg's images). Returns the list of loss.all values (python floats)."""
        n = len(batches) if n_steps is None else int(n_steps)
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_pipe", None) is None:
            self._pipe = dict(copy=torch.cuda.Stream(self.device), staging=[{}, {}],
                              h2d=[torch.cuda.Event(), torch.cuda.Event()], loaded=[torch.cuda.Event(), torch.cuda.Event()],
                              read=[torch.cuda.Event(), torch.cuda.Event()],
                              loss=[torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)])
        P = self._pipe

        def pairs(i):
            b = batches[i % len(batches)]
            out = [(self.var[k], t) for k, t in b.items()]
            return out + (list(extra(i)) if extra is not None else [])

        def stage(i, slot):
            P["copy"].wait_event(P["loaded"][slot])            # the step that last consumed this slot has taken its copy
            with torch.cuda.stream(P["copy"]), torch.no_grad():
                for dst, src in pairs(i):
                    buf = P["staging"][slot].get(id(dst))
                    if buf is None:
                        buf = P["staging"][slot][id(dst)] = torch.empty_like(dst.detach())
                    buf.copy_(src, non_blocking=True)
            P["h2d"][slot].record(P["copy"])

        losses = []
        for ev in P["loaded"]:
            ev.record(main)
        if n > 0:
            stage(0, 0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                stage(i + 1, slot ^ 1)
            main.wait_event(P["h2d"][slot])
            with torch.no_grad():
                for dst, _ in pairs(i):
                    dst.copy_(P["staging"][slot][id(dst)], non_blocking=True)
            P["loaded"][slot].record(main)
            loss = self()
            P["loss"][slot].copy_(loss["all"].detach().reshape(1), non_blocking=True)
            P["read"][slot].record(main)
            if i >= 1:
                P["read"][slot ^ 1].synchronize()
                losses.append(float(P["loss"][slot ^ 1]))
        if n > 0:
            P["read"][(n - 1) & 1].synchronize()
            losses.append(float(P["loss"][(n - 1) & 1]))
        return losses

    def _forward_backward(self):
        self.flat.zero()
        for k in GRAD_LEAVES:
            if k in self.var:
                self.var[k].grad = None
        main = torch.cuda.current_stream(self.device)
        if self.side_work is not None:
            if self._side is not None:
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self.side_work()
            else:
                self.side_work()
        # parameter gradients go straight into the flat buffer's views (render_fn.FUSED_GRAD_ACCUMULATION)
        old, render_fn.FUSED_GRAD_ACCUMULATION = render_fn.FUSED_GRAD_ACCUMULATION, True
        try:
            _, loss = self.graph(self.opt, self.var, training=True, get_loss=True)
            loss["all"].backward()
        finally:
            render_fn.FUSED_GRAD_ACCUMULATION = old
        if self.side_work is not None and self._side is not None:
            main.wait_stream(self._side)
        return loss

    def _capture(self, warmup):
        timers_were = rn.TIMERS.enabled
        rn.TIMERS.enabled = False
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):              # allocator / lazy-init warm-up on the capture stream
                self._forward_backward()
                self.flat.all_reduce()
                self.optim.step()
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        rn.TIMERS.enabled = timers_were
        rn.TIMERS.capturing = True
        n0 = rn.TIMERS.launches
        self._single = self.flat.world == 1 or os.environ.get("SC_ALLREDUCE_IN_GRAPH", "1") != "0"
        # NCCL's watchdog thread polls CUDA events while we capture: relaxed mode keeps its calls from invalidating the capture
        kw = dict(capture_error_mode="thread_local") if self.flat.world > 1 else {}
        self._g_main = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._g_main, **kw):
            self.loss = self._forward_backward()
            if self._single:
                self.flat.all_reduce()
                self.optim.step()
        if not self._single:
            self._g_optim = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._g_optim, pool=self._g_main.pool()):
                self.optim.step()
        rn.TIMERS.capturing = False
        self.launches_per_step = rn.TIMERS.launches - n0
        rn.TIMERS.launches = n0
        rn.invalidate_blob_cache()        # blobs cached during capture live in the graph's pool and go stale on replay

    # ------------------------------------------------------------------------------------------------------
    def __call__(self):
        """Run one step on the tensors in self.var; returns the loss dict (device scalars; 'all' is the total)."""
        if self._g_main is None:
            self.loss = self._forward_backward()
            self.flat.all_reduce()
            self.optim.step()
            return self.loss
        self._g_main.replay()
        if not self._single:
            self.flat.all_reduce()
            self._g_optim.replay()
        # a replayed optimiser step rewrites the weights without bumping their version counters: packed weight blobs cached
        # by eager callers (validation renders, eval_3D.compute_level_grid) would otherwise stay at the first call's weights
        rn.invalidate_blob_cache()
        rn.TIMERS.count(self.launches_per_step)
        return self.loss
