#!/usr/bin/env python
"""bench.py — train-step images/s of ShapeClipper's hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = (BASELINE.json configs[1]: batch 16 per GPU, 512 rays x 64 samples, K=5 neighbours, 1 neighbour view):
  CLIP ViT-B/32 encode of the batch images + cosine top-k against the bank (when built, see config.clip),
  render(query view) + render(CLIP-neighbour view) + eikonal queries, the seven render losses, backward
  (double backward through the SDF MLP), one flat gradient all-reduce (N > 1), Adam step.
`value` times the step with the batch already resident in HBM; `e2e` times it through the public API with host
(pinned) batches: host->device copies and the device->host read of the loss inside the timed region.
`--impl reference` times the CPU restatement of the reference path (oracle/, the reference is Python and does not
travel to the GPU box) on the host cores for a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train-step images/sec (render+CLIP+loss+bwd)"
UNIT = "images/s"
FLOP_FWD_PER_POINT = 199424.0       # SURVEY.md §8d convention: 2*(40320 SDF + 40320 grad-SDF + 19072 RGB)
FLOP_BWD_PER_POINT = 398848.0       # train fwd+bwd = 3x fwd  ->  backward kernel = 2x fwd
GRAPH_PARAMS = 36800589             # parameters of the reference Graph (flat all-reduce size)
RENDER_BWD_DRAM_BYTES = 2462139000  # dram__bytes_read.sum + dram__bytes_write.sum of one render_tc_bwd_kernel<0> launch at this shape
                                    # (ncu --set full, profiles/r01h_render_tc_bwd_ncu_summary.txt: 1.93 GB of saved activations read)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="images per GPU")
    ap.add_argument("--ref-batch", type=int, default=2, help="images per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-stream", action="store_true", help="run the CLIP leg on the main stream instead of a forked one")
    ap.add_argument("--eager", action="store_true", help="launch every kernel from the host instead of replaying CUDA graphs")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        time.sleep(0.05)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        mx = max(int(float(r[1])) for r in self.rows if len(r) >= 2 and r[1].replace(".", "").isdigit())
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=reasons, samples=len(sm))


# --------------------------------------------------------------------------------------------------- reference arm
def oracle_step_fn(opt, batch, seed=0):
    """One training step of the CPU restatement (oracle/): 2 renders + losses + backward + Adam."""
    import torch
    from oracle import render_ref as R, loss_ref
    from shapeclipper_b200.implicit import SDFNetwork, RGBNetwork
    torch.manual_seed(seed)
    sdf, rgb = SDFNetwork(opt), RGBNetwork(opt)
    beta = torch.tensor(0.1, requires_grad=True)
    sp = {k: v.detach().clone().requires_grad_(True) for k, v in sdf.state_dict().items()}
    rp = {k: v.detach().clone().requires_grad_(True) for k, v in rgb.state_dict().items()}
    params = list(sp.values()) + list(rp.values()) + [beta]
    optim = torch.optim.Adam(params, lr=1e-4)
    leaves = {k: batch[k].clone().requires_grad_(True) for k in ("pose", "intr", "scale_dist", "proj_latent_sdf", "proj_latent_rgb")}
    B = batch["rgb_input"].shape[0]

    def step():
        optim.zero_grad(set_to_none=True)
        out = R.render(sp, rp, beta, leaves["pose"], leaves["intr"], leaves["scale_dist"], leaves["proj_latent_sdf"],
                       leaves["proj_latent_rgb"], opt.H, opt.W, ray_idx=batch["ray_idx"], training=True)
        L = loss_ref.render_losses(out, batch["rgb_input"], batch["mask_input"],
                                   batch["normal_input"] @ leaves["pose"][..., :3], B)
        out2 = R.render(sp, rp, beta, batch["pose_NN"][..., 0], batch["intr_NN"][..., 0], batch["scale_dist_NN"][..., 0],
                        leaves["proj_latent_sdf"], batch["proj_latent_rgb_NN"][..., 0], opt.H, opt.W,
                        ray_idx=batch["ray_idx_NN"][..., 0], training=True)
        L2 = loss_ref.render_losses(out2, batch["rgb_input_NN"][..., 0], batch["mask_input_NN"][..., 0],
                                    batch["normal_input_NN"][..., 0] @ batch["pose_NN"][..., 0][..., :3], B)
        total = loss_ref.weighted_total(L) + 1.0 * L2["render"] + 0.5 * L2["mask"] + 0.01 * L2["normal"]
        total.backward()
        optim.step()
        return float(total)
    return step


def cpu_reference_rate(opt, images, steps, warmup):
    import torch
    from shapeclipper_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    batch = synthetic.make_batch(opt, images, seed=0, pin=False)
    step = oracle_step_fn(opt, batch)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return images / dt, dt, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from shapeclipper_b200 import options
    opt = options.default_options(device="cpu")
    steps, warm = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
    rate, dt, cores = cpu_reference_rate(opt, a.ref_batch, steps, warm)
    sample = "%d images/step of the same workload (512 rays x 64 samples, 2 renders, losses, backward, Adam), %d timed steps" % (a.ref_batch, steps)
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=a.gpus, steps=steps, warmup=warm, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=workload_config(a, opt, clip=False),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(a, opt, clip):
    return dict(workload="configs[1]: batch=%d/GPU Pix3D-shaped synthetic, %d rays x %d samples, k_nearest=%d, n_views=%d, "
                         "render+losses+backward+Adam%s" % (a.batch, int(opt.render.rand_sample), opt.render.n_samples_uniform,
                                                            opt.data.k_nearest, opt.reg.n_views, "+CLIP ViT-B/32" if clip else ""),
                per_gpu_batch=a.batch, image_size=[opt.H, opt.W], l2="inputs cycle through 4 distinct batches; the render kernels' "
                "working set per step (2 x 1.97 GB of saved activations + 97 MB of per-CTA scratch) exceeds the 126 MB L2", parallelism="dp%d" % a.gpus, clip=clip)


# --------------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as tdist
    from shapeclipper_b200 import _render_native as rn, dist as scdist, options, synthetic
    from shapeclipper_b200.graph import HotPathGraph
    from shapeclipper_b200.step import TrainStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback exists)")
    rank, world, local = scdist.setup()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    opt = options.default_options(device=str(dev))
    opt.reg.device_sampling = True                  # neighbour draw stays on the GPU (no host sync in the step)
    opt.render.device_rng = True                    # stratified jitter / eikonal samples from the CUDA generator
    torch.manual_seed(0)
    graph = HotPathGraph(opt).to(dev)
    params = list(graph.renderer.parameters())
    hot = sum(p.numel() for p in params)
    flat = scdist.FlatGradients(params, extra=(GRAPH_PARAMS - hot) if world > 1 else 0, device=dev)
    flat.broadcast_parameters()
    optim = torch.optim.Adam(params, lr=1e-4, foreach=True, capturable=not a.eager)
    clip_ctx = None
    try:
        from shapeclipper_b200 import clip as scclip
        clip_ctx = scclip.bench_context(opt, a.batch, dev)
    except ImportError:
        clip_ctx = None
    batches = [synthetic.make_batch(opt, a.batch, seed=1000 * rank + i) for i in range(4)]
    resident = [{k: t.to(dev) for k, t in b.items()} for b in batches]
    h2d_bytes = sum(t.numel() * t.element_size() for t in batches[0].values())
    if clip_ctx is not None:
        h2d_bytes += clip_ctx.h2d_bytes

    rn.TIMERS.reset()
    rn.TIMERS.enabled = True                        # kernel spans: CUDA events (external event nodes inside the graphs)
    step = TrainStep(opt, graph, optim, flat, batches[0], dev, side_work=(clip_ctx.run if clip_ctx is not None else None),
                     use_cuda_graph=not a.eager, side_stream=not a.no_side_stream)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return float(ms) / steps

    def resident_step(i):     # batch already in HBM: device->device refresh of the step's input tensors, then the step
        step.load(resident[i % len(resident)])
        if clip_ctx is not None:
            clip_ctx.static_images.copy_(clip_ctx.images[i % 2], non_blocking=True)
        return step()

    for i in range(max(3, a.warmup)):
        resident_step(i)
    # ---- device-resident throughput (value)
    rn.TIMERS.reset()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_step = timed(resident_step, a.steps)
    launches = rn.TIMERS.launches / a.steps
    # nvidia-smi samples every 20 ms: when the timed region is shorter than ~150 ms, every rank keeps the same load running
    # (untimed; the count follows from the max-reduced step time, so all ranks agree) until the sampler has seen enough of it
    extra_clock_steps = 0
    if a.steps * ms_step < 150.0:
        extra_clock_steps = int((150.0 - a.steps * ms_step) / max(ms_step, 1e-3)) + 1
        for i in range(extra_clock_steps):
            resident_step(i)
        barrier()
    clocks = sampler.stop() if sampler else None
    if clocks is not None and extra_clock_steps:
        clocks["untimed_steps_sampled_too"] = extra_clock_steps
    # per-kernel durations: eager = every launch of the timed region; graphs = the event nodes inside the graph, read after
    # the last timed replay and after 5 more replays (a synchronize between them)
    if a.eager:
        kernel_ms = rn.TIMERS.totals_ms()
        span_steps = a.steps + extra_clock_steps
    else:
        acc, span_steps = {}, 0
        for rep in range(6):
            if rep:
                resident_step(rep)
            for k, (ms, n) in rn.TIMERS.graph_ms().items():
                t = acc.setdefault(k, [0.0, 0])
                t[0] += ms
                t[1] += n
            span_steps += 1
        kernel_ms = {k: tuple(v) for k, v in acc.items()}
    rn.TIMERS.enabled = False

    # ---- end to end through the public API: pinned host batch -> device, step, loss back to the host
    def e2e_step(i):
        step.load(batches[i % len(batches)])
        if clip_ctx is not None:
            clip_ctx.static_images.copy_(clip_ctx.host_images[i % 2], non_blocking=True)
        loss = step()
        return float(loss["all"].detach())         # device -> host read of the result
    for i in range(3):
        e2e_step(i)
    ms_e2e = timed(e2e_step, a.steps)

    if rank != 0:
        scdist.cleanup()
        return
    pk = peaks()
    images = a.batch * world
    S, R = opt.render.n_samples_uniform, int(opt.render.rand_sample)
    pts = a.batch * R * S                                       # sample points per render launch
    bwd_ms, bwd_n = kernel_ms.get("render_bwd", (0.0, 1))
    fwd_ms, fwd_n = kernel_ms.get("render_fwd", (0.0, 1))
    bwd_avg = bwd_ms / max(bwd_n, 1)
    ffma_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    achieved = pts * FLOP_BWD_PER_POINT / (bwd_avg * 1e-3) / 1e12 if bwd_avg > 0 else 0.0
    share = lambda ms: ms / span_steps / ms_step
    roofline = dict(kernel="render_tc_bwd_kernel<0>", bound="tensor", achieved=achieved, peak=pk["bf16_sustained"], unit="TFLOP/s",
                    frac=achieved / pk["bf16_sustained"], traffic=RENDER_BWD_DRAM_BYTES,
                    peak_source=pk["source"] + " bf16 sustained (kernel timed inside the step)",
                    algorithmic_flops_per_launch=pts * FLOP_BWD_PER_POINT, avg_launch_ms=bwd_avg,
                    note="fp32-class products on tcgen05: every operand is a hi/lo bf16 pair and every product 3 MMAs (the 1e-4 "
                         "parity target rules out plain bf16/tf32 operands, BASELINE.md §2), so this scheme tops out at 1/3 of "
                         "the bf16 peak; flops follow SURVEY.md §8d's convention (2*MAC of the GEMMs as the reference runs them)",
                    timing="CUDA events around the kernel on its launch stream" + ("" if a.eager else
                           " (external event nodes inside the step's CUDA graph; 6 replays)"),
                    fp32_ffma_peak_tflops=ffma_peak, frac_of_fp32_ffma=achieved / ffma_peak,
                    share_of_step=dict(render_bwd=share(bwd_ms), render_fwd=share(fwd_ms),
                                       **{k: share(v[0]) for k, v in kernel_ms.items() if k.startswith("sdf") or k.startswith("clip")}),
                    share_note=("clip_encode runs on a forked low-priority stream and fills the SMs the small kernels between the render "
                                "launches leave idle: its share is the span it is spread over, not exclusive time") if not a.no_side_stream else None,
                    render_fwd_tflops=(pts * FLOP_FWD_PER_POINT / (fwd_ms / max(fwd_n, 1) * 1e-3) / 1e12) if fwd_ms > 0 else None)
    line = dict(metric=METRIC, value=images / (ms_step * 1e-3), unit=UNIT, n_gpus=world, steps=a.steps, warmup=max(3, a.warmup),
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(a, opt, clip=clip_ctx is not None), impl="ours",
                e2e=dict(value=images / (ms_e2e * 1e-3), unit=UNIT, ms_per_step=ms_e2e, h2d_bytes_per_step=h2d_bytes,
                         d2h_bytes_per_step=4),
                gpu_launches=launches, clocks=clocks, roofline=roofline)
    line["config"]["execution"] = "eager launches" if a.eager else "CUDA graphs (forward+losses+backward, optimiser) replayed per step"
    if world == 1 and not a.no_cpu_baseline:
        rate, dt, cores = cpu_reference_rate(options.default_options(device="cpu"), 1, 1, 1)
        line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=cores, kind="port",
                                    sample="1 image/step of the same workload (512 rays x 64 samples, 2 renders, losses, "
                                           "backward, Adam; CLIP excluded), 1 warm-up + 1 timed step of oracle/")
    print(json.dumps(line), flush=True)
    scdist.cleanup()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
