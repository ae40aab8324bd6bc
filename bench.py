#!/usr/bin/env python
"""bench.py — train-step images/s of ShapeClipper's hot path on N B200s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = (per-GPU shard of BASELINE.json configs[3]: batch 32 per GPU — 256 global on 8 GPUs — 512 rays x 64 samples, K=5
neighbours, 1 neighbour view; `--batch 16` is configs[1], which the N=1 line also carries as a sub-record):
  CLIP ViT-B/32 encode of the batch images + cosine top-k against the bank (when built, see config.clip),
  render(query view) + render(CLIP-neighbour view) + eikonal queries, the seven render losses, backward
  (double backward through the SDF MLP), one flat gradient all-reduce (N > 1), Adam step.
`value` times the step with the batch already resident in HBM; `e2e` times it through the public API with host
(pinned) batches: host->device copies and the device->host read of the loss inside the timed region.
At N=1 the JSON line also carries `configs`: driver-run sub-records for the other BASELINE.json configurations —
configs[0] (CPU reference path), configs[1] (batch 16 step), configs[2] (batch 64: CLIP encode + top-k, and the two-render
128x128 TRAINING step in recompute mode) and configs[4] (evaluate.py path: vox_res=100 level grid + chamfer 100k x 100k next to
the recompiled reference kernel) — each with its own roofline. At N>1 `allreduce` times the flat gradient all-reduce alone.
`--impl reference` times the reference's OWN Python modules (model/renderer.py, model/implicit.py, model/loss.py, run from the
bytecode oracle/build_ref.py staged under oracle/_ref/py; the oracle port when that is absent) on the host cores for a
bounded sample of the same workload, CLIP leg included (architecture twin, oracle/clip_ref.py — openai/CLIP is not available).
"""
import argparse
import json
import math
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
# dram__bytes_read.sum + dram__bytes_write.sum of one render_tc_bwd_kernel<0, 0, 1> launch (ncu --set full, 512 rays x 64 samples,
# profiles/r02n_render_tc_b{16,32}_ncu_summary.txt), by images per launch: the saved activations read back (3.75 KB per sample point)
# plus the L2 write-back of the per-CTA scratch planes
RENDER_BWD_DRAM_BYTES = {16: 1931685000 + 393357000, 32: 3863811000 + 762239000}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (32 x 8 GPUs = BASELINE configs[3]; 16 = configs[1])")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0]/[1]/[2]/[4] sub-records of the N=1 line")
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
CLIP_GFLOP_PER_IMAGE = 8.82          # ViT-B/32 image tower, 2*MAC
SDF_FLOP_PER_POINT = 80640.0         # SDF value only (level grid)
CHAMFER_PAIR_PEAK = 148 * 128 * 1.965e9 / 6.0     # <= 6 FMA-pipe lane-ops per pair: 6.3 T pairs/s


def _staged_reference():
    """The reference's own modules from oracle/_ref/py (bytecode staged by oracle/build_ref.py), or None."""
    try:
        from oracle import build_ref
        if not build_ref.staged_available():
            return None
        import types
        for name in ("vigra", "mcubes", "trimesh", "seaborn", "chamfer_3D"):
            sys.modules.setdefault(name, types.ModuleType(name))
        tc = types.ModuleType("termcolor")
        tc.colored = lambda s_, *a, **k: str(s_)
        sys.modules.setdefault("termcolor", tc)
        build_ref.install_staged_importer()
        import importlib
        return dict(implicit=importlib.import_module("model.implicit"), renderer=importlib.import_module("model.renderer"),
                    loss=importlib.import_module("model.loss"), camera=importlib.import_module("utils.camera"),
                    util=importlib.import_module("utils.util"), opt_json=os.path.join(build_ref.PY_OUT, "options", "pix3d", "config.json"))
    except Exception as ex:  # noqa: BLE001
        print("bench.py: staged reference modules unusable (%r): timing the oracle port instead" % (ex,), file=sys.stderr)
        return None


def reference_step_fn(mods, batch, H, W, n_rays, seed=0, n_samples=64):
    """One training step of the hot path with the REFERENCE's own classes (model/renderer.py Renderer, model/implicit.py
    SDFNetwork / RGBNetwork, model/loss.py Loss, utils/camera.py transform_normal): 2 renders + 7 losses + backward + Adam.
    The CNN encoders (out of scope for both arms) are replaced by the batch's latent codes / poses, as in our arm."""
    import torch
    opt = mods["util"].EasyDict(json.load(open(mods["opt_json"])))
    opt.device, opt.H, opt.W = "cpu", H, W
    opt.render.rand_sample = n_rays
    opt.render.n_samples_uniform = n_samples
    torch.manual_seed(seed)
    sdf, rgb = mods["implicit"].SDFNetwork(opt), mods["implicit"].RGBNetwork(opt)
    ren = mods["renderer"].Renderer(opt, sdf, rgb)
    fns = mods["loss"].Loss(opt)
    optim = torch.optim.Adam(ren.parameters(), lr=1e-4)
    leaves = {k: batch[k].clone().requires_grad_(True) for k in ("pose", "intr", "scale_dist", "proj_latent_sdf", "proj_latent_rgb")}
    B = batch["rgb_input"].shape[0]
    tn = mods["camera"].transform_normal
    lw = opt.loss_weight

    def losses(out, rgb_t, mask_t, normal_t, eik):
        rgb_o, mask_o, _, _, normal_o, g_eik = out
        valid = (mask_t > 0.5) & (mask_o > 0.5)
        L = [fns.MSE_loss(rgb_o, rgb_t), fns.mask_loss(mask_o, mask_t), fns.normal_loss(normal_o, normal_t, valid, tolerance=opt.reg.normal_tol)]
        return L + ([fns.MSE_loss(g_eik.view(B, -1), 1)] if eik else [])

    def step():
        optim.zero_grad()
        out = ren(opt, leaves["pose"], leaves["intr"], leaves["scale_dist"], leaves["proj_latent_sdf"], leaves["proj_latent_rgb"],
                  ray_idx=batch["ray_idx"], training=True)
        L = losses(out, batch["rgb_input"], batch["mask_input"], tn(batch["normal_input"], leaves["pose"]), True)
        pose_n = batch["pose_NN"][..., 0]
        out2 = ren(opt, pose_n, batch["intr_NN"][..., 0], batch["scale_dist_NN"][..., 0], leaves["proj_latent_sdf"],
                   batch["proj_latent_rgb_NN"][..., 0], ray_idx=batch["ray_idx_NN"][..., 0], training=True)
        L2 = losses(out2, batch["rgb_input_NN"][..., 0], batch["mask_input_NN"][..., 0], tn(batch["normal_input_NN"][..., 0], pose_n), False)
        total = (lw.render * L[0] + lw.mask * L[1] + lw.normal * L[2] + lw.eikonal * L[3]
                 + lw.nearest_img * L2[0] + lw.nearest_mask * L2[1] + lw.nearest_normal * L2[2])
        total.backward()
        optim.step()
        return float(total)
    return step


def oracle_step_fn(opt, batch, seed=0):
    """The same step on the CPU restatement (oracle/): used where the staged reference modules are absent, and by the tests."""
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


def clip_cpu_fn(images, bank_size=4096):
    """CLIP leg on the CPU: ViT-B/32 image tower (oracle/clip_ref.py, the fp32 architecture twin — openai/CLIP itself is not
    installable offline) on `images` images + cosine top-6 against the bank, as NN_annotator.calc_matches does per query."""
    import torch
    from oracle import clip_ref
    cfg = clip_ref.CONFIGS["ViT-B/32"]
    p = clip_ref.random_params(cfg, seed=0)
    g = torch.Generator().manual_seed(7)
    img = torch.randn(images, 3, 224, 224, generator=g)
    bank = torch.nn.functional.normalize(torch.randn(bank_size, cfg["out_dim"], generator=g), dim=-1)

    def run():
        with torch.no_grad():
            e = torch.nn.functional.normalize(clip_ref.encode_image(p, cfg, img).float(), dim=-1)
            return (e @ bank.t()).topk(6, dim=-1)
    return run


def cpu_reference_rate(images, steps, warmup, clip=True, H=224, W=224, n_rays=512):
    """-> (images/s, seconds per step, threads, kind): the reference's own modules when staged, else the oracle port."""
    import torch
    from shapeclipper_b200 import options, synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    opt = options.default_options(H=H, W=W, device="cpu")
    opt.render.rand_sample = n_rays
    batch = synthetic.make_batch(opt, images, seed=0, pin=False)
    mods = _staged_reference()
    step = reference_step_fn(mods, batch, H, W, n_rays) if mods is not None else oracle_step_fn(opt, batch)
    clip_run = clip_cpu_fn(images) if clip else (lambda: None)
    for _ in range(warmup):
        clip_run(); step()
    t0 = time.perf_counter()
    for _ in range(steps):
        clip_run(); step()
    dt = (time.perf_counter() - t0) / steps
    return images / dt, dt, torch.get_num_threads(), ("reference" if mods is not None else "port")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from shapeclipper_b200 import options
    opt = options.default_options(device="cpu")
    steps, warm = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
    rate, dt, cores, kind = cpu_reference_rate(a.ref_batch, steps, warm)
    sample = ("%d images/step of the same workload (CLIP ViT-B/32 encode + top-6 of a 4096 bank, 512 rays x 64 samples, 2 renders, 7 losses, "
              "backward, Adam), %d timed steps; %s" % (a.ref_batch, steps, "the reference's own model/renderer.py + model/implicit.py + "
              "model/loss.py on the host cores" if kind == "reference" else "oracle/ port (staged reference modules absent)"))
    line = dict(metric=METRIC, value=rate, unit=UNIT, n_gpus=a.gpus, steps=steps, warmup=warm, ms_per_step=dt * 1e3,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=workload_config(a.batch, a.gpus, opt, clip=True),
                cpu_baseline=dict(value=rate, unit=UNIT, cores=cores, kind=kind, sample=sample),
                e2e=dict(value=rate, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def workload_config(batch, gpus, opt, clip, name=None):
    rays = int(opt.render.rand_sample) if opt.render.rand_sample else opt.H * opt.W
    if name is None:
        name = "configs[1]" if batch == 16 else ("per-GPU shard of configs[3] (256 global on 8 GPUs)" if batch == 32 else "configs[1] shape at batch %d" % batch)
    return dict(workload="%s: batch=%d/GPU Pix3D-shaped synthetic, %d rays x %d samples, k_nearest=%d, n_views=%d, "
                         "render+losses+backward+Adam%s" % (name, batch, rays, opt.render.n_samples_uniform,
                                                            opt.data.k_nearest, opt.reg.n_views, "+CLIP ViT-B/32" if clip else ""),
                per_gpu_batch=batch, image_size=[opt.H, opt.W], l2="inputs cycle through 4 distinct batches; the render kernels' "
                "working set per step (saved activations, 3.75 KB per sample point, + 97 MB of per-CTA scratch) exceeds the 126 MB L2",
                parallelism="dp%d" % gpus, clip=clip)


# --------------------------------------------------------------------------------------------------- our arm
def measure_train_step(a, opt, batch_size, steps, warmup, rank, world, dev, clip=True, sample_clocks=True, n_batches=4):
    """Builds the step (HotPathGraph + TrainStep + CLIP leg) for `batch_size` images per GPU and measures it: device-resident
    throughput, per-kernel spans, end-to-end throughput from pinned host batches. Returns a dict (rank 0 reads it)."""
    import torch
    import torch.distributed as tdist
    from shapeclipper_b200 import _render_native as rn, dist as scdist, synthetic
    from shapeclipper_b200.graph import HotPathGraph
    from shapeclipper_b200.step import TrainStep

    torch.manual_seed(0)
    graph = HotPathGraph(opt).to(dev)
    params = list(graph.renderer.parameters())
    hot = sum(p.numel() for p in params)
    flat = scdist.FlatGradients(params, extra=(GRAPH_PARAMS - hot) if world > 1 else 0, device=dev)
    flat.broadcast_parameters()
    optim = torch.optim.Adam(params, lr=1e-4, foreach=True, capturable=not a.eager)
    clip_ctx = None
    if clip:
        from shapeclipper_b200 import clip as scclip
        # inside the step the tower runs as cooperative launches of 2 phases each: ~0.1 ms pieces that the render stream's kernels
        # interleave with (one 1.6-3 ms cooperative launch holds every SM; measured 12.31 -> 12.01 ms/step at batch 32)
        clip_ctx = scclip.bench_context(opt, batch_size, dev, precision=os.environ.get("SC_BENCH_CLIP", "split"),
                                        launch_group=int(os.environ.get("SC_TOWER_GROUP", "2")) if not a.no_side_stream else 0)
    batches = [synthetic.make_batch(opt, batch_size, seed=1000 * rank + i) for i in range(n_batches)]
    resident = [{k: t.to(dev) for k, t in b.items()} for b in batches]
    h2d_bytes = sum(t.numel() * t.element_size() for t in batches[0].values())
    if clip_ctx is not None:
        h2d_bytes += clip_ctx.h2d_bytes

    rn.TIMERS.reset()
    rn.TIMERS.graph_events = {}
    rn.TIMERS.enabled = True                        # kernel spans: CUDA events (external event nodes inside the graphs)
    step = TrainStep(opt, graph, optim, flat, batches[0], dev, side_work=(clip_ctx.run if clip_ctx is not None else None),
                     use_cuda_graph=not a.eager, side_stream=not a.no_side_stream)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(n):
            fn(i)
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
        return float(ms) / n

    def resident_step(i):     # batch already in HBM: device->device refresh of the step's input tensors, then the step
        step.load(resident[i % len(resident)])
        if clip_ctx is not None:
            clip_ctx.static_images.copy_(clip_ctx.images[i % 2], non_blocking=True)
        return step()

    for i in range(max(3, warmup)):
        resident_step(i)
    rn.TIMERS.reset()
    sampler = ClockSampler(dev.index) if (rank == 0 and sample_clocks) else None
    if sampler:
        sampler.start()
    ms_step = timed(resident_step, steps)
    launches = rn.TIMERS.launches / steps
    # nvidia-smi samples every 20 ms: when the timed region is shorter than ~150 ms, every rank keeps the same load running
    # (untimed; the count follows from the max-reduced step time, so all ranks agree) until the sampler has seen enough of it
    extra_clock_steps = 0
    if sample_clocks and steps * ms_step < 150.0:
        extra_clock_steps = int((150.0 - steps * ms_step) / max(ms_step, 1e-3)) + 1
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
        span_steps = steps + extra_clock_steps
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

    # ---- end to end through the public API: pinned host batch -> device, step, loss back to the host, EVERY step.
    # (a) serial: copy, step, read the loss, one after the other; (b) TrainStep.run_epoch, the loop a prefetching loader gives
    # (model/runner.py:198-225): the same copies and reads, batch i+1 in flight while step i runs. (b) is the e2e figure.
    def e2e_step(i):
        step.load(batches[i % len(batches)])
        if clip_ctx is not None:
            clip_ctx.static_images.copy_(clip_ctx.host_images[i % 2], non_blocking=True)
        loss = step()
        return float(loss["all"].detach())         # device -> host read of the result
    for i in range(3):
        e2e_step(i)
    ms_e2e_serial = timed(e2e_step, steps)
    extra = (lambda i: [(clip_ctx.static_images, clip_ctx.host_images[i % 2])]) if clip_ctx is not None else None
    step.run_epoch(batches, 3, extra=extra)
    e2e_losses = []
    ms_e2e = timed(lambda i: e2e_losses.extend(step.run_epoch(batches, steps, extra=extra)), 1) / steps
    e2e_ok = len(e2e_losses) == steps and all(math.isfinite(v) for v in e2e_losses)       # one finite loss read back per step

    # ---- the flat gradient all-reduce alone (N > 1): algorithm bandwidth and ring bus bandwidth 2 (N-1)/N x bytes / t
    allreduce = None
    if world > 1:
        for _ in range(3):
            flat.all_reduce()
        ms_ar = timed(lambda i: flat.all_reduce(), 20)
        nbytes = flat.flat.numel() * 4
        allreduce = dict(bytes=nbytes, ms=ms_ar, algbw_gbs=nbytes / ms_ar / 1e6, busbw_gbs=2.0 * (world - 1) / world * nbytes / ms_ar / 1e6,
                         guide_busbw_gbs=725.0, in_step_graph=bool(getattr(step, "_single", False)),
                         note="one NCCL all-reduce (op AVG: the 1/world scale is folded into the reduction) of the flat fp32 gradient buffer "
                              "(36 800 589 floats, the reference Graph's parameter count); inside the step's CUDA graph")

    pk = peaks()
    S, R = opt.render.n_samples_uniform, (int(opt.render.rand_sample) if opt.render.rand_sample else opt.H * opt.W)
    pts = batch_size * R * S                                    # sample points per render launch
    bwd_ms, bwd_n = kernel_ms.get("render_bwd", (0.0, 1))
    fwd_ms, fwd_n = kernel_ms.get("render_fwd", (0.0, 1))
    refwd_ms, refwd_n = kernel_ms.get("render_fwd_for_backward", (0.0, 0))      # chunked backward: the forward run again per chunk
    # per RENDER (one forward launch each); a chunked backward is several launches per render: their sum
    bwd_avg, fwd_avg = bwd_ms / max(fwd_n, 1), fwd_ms / max(fwd_n, 1)
    ffma_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    achieved = pts * FLOP_BWD_PER_POINT / (bwd_avg * 1e-3) / 1e12 if bwd_avg > 0 else 0.0
    share = lambda ms: ms / span_steps / ms_step
    saved = rn.saved_buffer_bytes(batch_size, R, S)
    roofline = dict(kernel="render_tc_bwd_kernel<0>", bound="tensor", achieved=achieved, peak=pk["bf16_sustained"], unit="TFLOP/s",
                    frac=achieved / pk["bf16_sustained"],
                    traffic=(RENDER_BWD_DRAM_BYTES.get(batch_size) if (R == 512 and S == 64 and saved) else None),
                    traffic_note="dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape (ncu --set full, profiles/r02n_render_tc_b*_ncu_summary.txt; "
                                 "null at shapes that were not captured): the activation planes the forward saved (3.75 KB per sample point) read back once - "
                                 "the alternative, recomputing the forward per tile, needs no HBM but 19 more GEMM phases per tile and is slower (bench configs[2] runs it)",
                    peak_source=pk["source"] + " bf16 sustained (kernel timed inside the step)",
                    algorithmic_flops_per_launch=pts * FLOP_BWD_PER_POINT, avg_launch_ms=bwd_avg,
                    launches_per_render=bwd_n / max(fwd_n, 1),
                    note="fp32-class products on tcgen05: every operand is a hi/lo bf16 pair and every product 3 MMAs (the 1e-4 "
                         "parity target rules out plain bf16/tf32 operands, BASELINE.md §2), so this scheme tops out at 1/3 of "
                         "the bf16 peak; flops follow SURVEY.md §8d's convention (2*MAC of the GEMMs as the reference runs them)",
                    timing="CUDA events around the kernel on its launch stream" + ("" if a.eager else
                           " (external event nodes inside the step's CUDA graph; 6 replays)"),
                    fp32_ffma_peak_tflops=ffma_peak, frac_of_fp32_ffma=achieved / ffma_peak,
                    backward_mode=("saved activations" if saved else
                                   ("chunked: per %d images the forward again with the saved-activation buffer (%.1f ms per render, not in "
                                    "avg_launch_ms) + the saved-activation backward; %d backward launches per render"
                                    % (max(1, round(batch_size * fwd_n / max(bwd_n, 1))), refwd_ms / max(fwd_n, 1), round(bwd_n / max(fwd_n, 1))))
                                   if refwd_n else "recompute per tile (saved planes would not fit)"),
                    share_of_step=dict(render_bwd=share(bwd_ms), render_fwd=share(fwd_ms), render_fwd_for_backward=share(refwd_ms),
                                       **{k: share(v[0]) for k, v in kernel_ms.items() if k.startswith("sdf") or k.startswith("clip")}),
                    share_note=("clip_encode runs on a forked low-priority stream and fills the SMs the small kernels between the render "
                                "launches leave idle: its share is the span it is spread over, not exclusive time") if (clip and not a.no_side_stream) else None,
                    render_fwd=dict(kernel="render_tc_fwd_kernel<0>", avg_launch_ms=fwd_avg,
                                    achieved=(pts * FLOP_FWD_PER_POINT / (fwd_avg * 1e-3) / 1e12) if fwd_avg > 0 else None,
                                    frac=(pts * FLOP_FWD_PER_POINT / (fwd_avg * 1e-3) / 1e12 / pk["bf16_sustained"]) if fwd_avg > 0 else None,
                                    unit="TFLOP/s"))
    images = batch_size * world
    res = dict(value=images / (ms_step * 1e-3), ms_per_step=ms_step,
               e2e=dict(value=images / (ms_e2e * 1e-3), unit=UNIT, ms_per_step=ms_e2e, h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=4,
                        losses_read=len(e2e_losses), losses_finite=e2e_ok,
                        api="TrainStep.run_epoch: every step's batch copied from pinned host memory and every step's loss read on the host, "
                            "batch i+1 in flight while step i runs",
                        serial=dict(value=images / (ms_e2e_serial * 1e-3), ms_per_step=ms_e2e_serial,
                                    note="the same copies and reads strictly one after the other (load, step, float(loss))")),
               gpu_launches=launches, clocks=clocks, roofline=roofline, allreduce=allreduce, clip=clip_ctx is not None)
    del step, graph, optim, flat, clip_ctx, resident, batches
    rn.invalidate_blob_cache()
    torch.cuda.empty_cache()
    return res


def _time_cuda(fn, iters, warm=3):
    import torch
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


def config2_record(a, dev, pk):
    """BASELINE configs[2]: batch 64 — CLIP ViT-B/32 encode + cosine top-6 (k_nearest = 5 after dropping self) in every precision
    mode the encoder has, and the two-render 128 x 128 TRAINING step (full-grid rays: 67 108 864 sample points per render; the
    saved planes of a whole render would be 252 GB, so the backward walks the batch in chunks of images: forward again with the
    saved-activation buffer, then the saved-activation backward; SC_RENDER_CHUNKED_BACKWARD=0 selects the per-tile recompute kernel)."""
    import torch
    from shapeclipper_b200 import clip as scclip, options
    rec = {}
    B = 64
    clip_rec = {}
    for prec in scclip.PRECISIONS:
        ctx = scclip.bench_context(None, B, dev, precision=prec)
        ms = _time_cuda(lambda: ctx.run(ctx.images[0]), 20)
        tf = CLIP_GFLOP_PER_IMAGE * B / ms                       # GFLOP / ms = TFLOP/s
        clip_rec[prec] = dict(ms_per_batch=ms, images_per_s=B / ms * 1e3, launches_per_encode=ctx.launches,
                              roofline=dict(bound="tensor", achieved=tf, peak=pk["bf16_sustained"], unit="TFLOP/s", frac=tf / pk["bf16_sustained"],
                                            note=scclip.PRECISION_NOTES.get(prec, ""), algorithmic_gflop_per_image=CLIP_GFLOP_PER_IMAGE,
                                            tensor_pipe_pct="see profiles/ (ncu sm__pipe_tensor_cycles_active per GEMM)"))
        del ctx
        torch.cuda.empty_cache()
    rec["clip_encode_topk_batch64"] = clip_rec
    # ---- 128 x 128 two-render training step, batch 64, full-grid rays
    opt = options.default_options(H=128, W=128, device=str(dev))
    opt.render.rand_sample = None
    opt.reg.device_sampling = True
    opt.render.device_rng = True
    steps = {}
    from shapeclipper_b200 import render_fn
    for mode in render_fn.STEP_PRECISIONS:
        render_fn.set_precision(forward=mode, backward=mode)
        try:
            r = measure_train_step(a, opt, B, steps=3, warmup=1, rank=0, world=1, dev=dev, clip=False, sample_clocks=False, n_batches=1)
        finally:
            render_fn.set_precision(forward="tc", backward="tc")
        steps[mode] = dict(images_per_s=r["value"], ms_per_step=r["ms_per_step"], e2e_images_per_s=r["e2e"]["value"],
                           sample_points_per_render=B * 128 * 128 * 64, roofline=r["roofline"], gpu_launches=r["gpu_launches"])
    rec["render_128x128_train_step_batch64"] = steps
    rec["workload"] = "configs[2]: batch=64, 128x128 full-grid render (2 renders, losses, backward, Adam) + CLIP-NN k=5 (encode + top-6 of a 4096 bank)"
    return rec


def config4_record(a, dev, pk, shapes=8):
    """BASELINE configs[4], the evaluate.py path per shape (eval.batch_size = 1, utils/eval_3D.py:52-103): SDF level grid at
    vox_res = 100 (1 030 301 queries) -> marching cubes -> 100 000 area-weighted surface samples -> chamfer against a 100 000-point
    ground-truth cloud -> F-score, everything on the device. The recompiled reference chamfer kernel (oracle/_ref) is timed next
    to ours on the same box; the reference's own mesh leg (PyMCubes + trimesh on the CPU) cannot be timed, both packages are absent."""
    import torch
    from shapeclipper_b200 import chamfer_3D, eval_3D, options
    from shapeclipper_b200.implicit import SDFNetwork
    opt = options.default_options(device=str(dev))
    opt.eval.vox_res = 100
    torch.manual_seed(0)
    sdf = SDFNetwork(opt).to(dev)
    var = options.Options(idx=torch.arange(1))
    pts = eval_3D.get_dense_3D_grid(opt, var).contiguous()
    g = torch.Generator().manual_seed(3)
    z = (torch.randn(shapes, 1, 64, generator=g) * 0.3).to(dev)
    N = opt.eval.num_points
    clouds = [torch.nn.functional.normalize(torch.randn(2, 1, N, 3, generator=g), dim=-1).mul(0.5).to(dev) for _ in range(2)]
    outs = [torch.zeros(1, N, device=dev), torch.zeros(1, N, device=dev),
            torch.zeros(1, N, dtype=torch.int32, device=dev), torch.zeros(1, N, dtype=torch.int32, device=dev)]
    i = [0]

    def grid():
        i[0] += 1
        return eval_3D.compute_level_grid(opt, sdf, z[i[0] % shapes], pts)

    def chamfer():
        i[0] += 1
        c = clouds[i[0] % 2]
        return chamfer_3D.forward(c[0], c[1], *outs)

    gen = torch.Generator(device=dev).manual_seed(0)
    level0 = grid()

    def mesh():
        return eval_3D.convert_to_explicit(opt, level0, 0., to_pointcloud=True, generator=gen)

    def shape():
        level = grid()
        _, pred = eval_3D.convert_to_explicit(opt, level, 0., to_pointcloud=True, generator=gen)
        i[0] += 1
        d1, d2, _, _ = eval_3D.chamfer_distance(opt, eval_3D.normalize_pc(pred), eval_3D.normalize_pc(clouds[i[0] % 2][1]))
        return eval_3D.compute_fscore(d1, d2, opt.eval.f_thresholds)
    ms_grid, ms_ch, ms_shape = _time_cuda(grid, shapes), _time_cuda(chamfer, shapes), _time_cuda(shape, shapes)
    ms_mesh = _time_cuda(mesh, shapes)
    n_tris = int(mesh()[0][0].triangles.shape[0])
    n_pts = 101 ** 3
    pairs = 2.0 * N * N
    rec = dict(workload="configs[4]: evaluate.py path per shape, vox_res=100 level grid + marching cubes + 100000 surface samples + "
                        "chamfer3D 100000 x 100000 + F-score, B=1, all on the device", shapes_timed=shapes,
               mesh=dict(ms=ms_mesh, triangles=n_tris, note="sc_mc_count + scan + sc_mc_emit + sc_tri_area + scan + searchsorted + sc_tri_sample; "
                         "HBM-bound byte work: 2 x 4.1 MB of grid reads + 36 B per triangle",
                         roofline=dict(bound="hbm", unit="GB/s", achieved=(2 * 101 ** 3 * 4 + 36 * n_tris + 12 * N) / ms_mesh / 1e6, peak=pk["hbm_gbs"],
                                       frac=(2 * 101 ** 3 * 4 + 36 * n_tris + 12 * N) / ms_mesh / 1e6 / pk["hbm_gbs"],
                                       note="launch- and sync-bound at this size (one host read of the triangle count per shape); not a bandwidth kernel in practice")),
               shapes_per_s=1e3 / ms_shape, ms_per_shape=ms_shape,
               level_grid=dict(ms=ms_grid, points=n_pts, roofline=dict(bound="tensor", unit="TFLOP/s", achieved=n_pts * SDF_FLOP_PER_POINT / ms_grid / 1e9,
                                                                     peak=pk["bf16_sustained"], frac=n_pts * SDF_FLOP_PER_POINT / ms_grid / 1e9 / pk["bf16_sustained"],
                                                                     note="3-MMA split operands: ceiling 1/3 of the bf16 peak")),
               chamfer=dict(ms=ms_ch, pairs=pairs, roofline=dict(bound="fp32 FMA pipe", unit="T pairs/s", achieved=pairs / ms_ch / 1e9,
                                                                 peak=CHAMFER_PAIR_PEAK / 1e12, frac=pairs / ms_ch / 1e9 / (CHAMFER_PAIR_PEAK / 1e12),
                                                                 hbm_gbs=(2 * N) * 20 / ms_ch / 1e6, hbm_peak_gbs=pk["hbm_gbs"],
                                                                 note="6 FMA-pipe lane-ops per pair at 1.965 GHz; algorithmic bytes 20 B/point, HBM irrelevant")))
    try:
        from oracle import build_ref                              # the kernel to beat: baseline leg only
        ref = build_ref.load()
        if ref is not None:
            ms_ref = _time_cuda(lambda: ref.forward(clouds[0][0], clouds[0][1], *outs), 3, warm=1)
            rec["chamfer"]["reference_kernel_ms"] = ms_ref
            rec["chamfer"]["speedup_vs_reference_kernel"] = ms_ref / ms_ch
            rec["chamfer"]["reference_kernel"] = "external/chamfer3D recompiled for sm_100a (oracle/_ref/chamfer_3D_ref.so), same box"
    except Exception as ex:  # noqa: BLE001
        rec["chamfer"]["reference_kernel_error"] = repr(ex)
    return rec


def ray_sampler_record(a, dev, pk, batch=32, H=224, W=224, n=512):
    """SURVEY §8f-4, the DataLoader's boundary-distance ray sampler (utils/util.py:237-248) for one batch of masks: the GPU transform
    + device draw next to the CPU transform the reference's workers run per image (vigra is absent: scipy's exact EDT, 1 thread)."""
    import numpy as np
    import torch
    from oracle import sampling_ref
    from shapeclipper_b200 import sampling
    g = np.random.RandomState(0)
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.stack([((yy - g.randint(60, 160)) ** 2 / float(g.randint(30, 80)) ** 2 + (xx - g.randint(60, 160)) ** 2 / float(g.randint(30, 80)) ** 2 < 1)
                  for _ in range(batch)]).astype(np.float32)
    md = torch.from_numpy(m).to(dev)
    ms_dist = _time_cuda(lambda: sampling.boundary_distance(md), 50)
    ms_draw = _time_cuda(lambda: sampling.sample_rays(md, n, 3.0), 50)
    t0 = time.perf_counter()
    for b in range(8):
        sampling_ref.boundary_distance_scipy(m[b] > 0.5)
    cpu_ms = (time.perf_counter() - t0) / 8 * 1e3
    bytes_alg = batch * H * W * 12.0                       # mask read by both passes + distance written, fp32
    return dict(workload="ray sampler: %d masks of %dx%d -> boundary distance transform + %d rays per image without replacement" % (batch, H, W, n),
                launches=2, transform_ms_per_batch=ms_dist, transform_plus_draw_ms_per_batch=ms_draw,
                cpu_reference=dict(ms_per_image=cpu_ms, cores=1, kind="port",
                                   note="scipy's exact EDT standing in for vigra.filters.boundaryDistanceTransform (absent), per image as in the DataLoader workers"),
                speedup_per_image_vs_one_core=cpu_ms * batch / ms_dist,
                roofline=dict(bound="hbm", unit="GB/s", achieved=bytes_alg / ms_dist / 1e6, peak=pk["hbm_gbs"], frac=bytes_alg / ms_dist / 1e6 / pk["hbm_gbs"],
                              note="12 B per pixel algorithmic; the column pass does H = 224 integer min-steps per pixel out of L2 (4.5 MB of "
                                   "uint16 row distances per batch): latency- and ALU-bound at this size, not HBM-bound"))


def config0_record():
    """BASELINE configs[0]: one 224 x 224 synthetic image, 32 x 32 rays x 32 samples, CLIP ViT-B/32, one forward + loss + backward
    on the CPU (the reference path itself, no GPU)."""
    import torch
    from shapeclipper_b200 import options, synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    mods = _staged_reference()
    opt = options.default_options(H=32, W=32, device="cpu")
    opt.render.rand_sample = 1024
    opt.render.n_samples_uniform = 32
    batch = synthetic.make_batch(opt, 1, seed=0, pin=False)
    if mods is not None:
        step = reference_step_fn(mods, batch, 32, 32, 1024, n_samples=32)
        kind = "reference"
    else:
        step, kind = oracle_step_fn(opt, batch), "port"
    clip_run = clip_cpu_fn(1)
    clip_run(); step()
    t0 = time.perf_counter(); clip_run(); t1 = time.perf_counter(); step(); t2 = time.perf_counter()
    return dict(workload="configs[0]: single 224x224 image, 32x32 rays x 32 samples (the 32^3 ray grid), CLIP ViT-B/32 forward + render x2 + "
                         "losses + backward + Adam on the host CPU", kind=kind, cores=torch.get_num_threads(), clip_encode_s=t1 - t0,
                render_step_s=t2 - t1, images_per_s=1.0 / (t2 - t0))


def run_ours(a):
    import torch
    from shapeclipper_b200 import dist as scdist, options

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback exists)")
    rank, world, local = scdist.setup()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    opt = options.default_options(device=str(dev))
    opt.reg.device_sampling = True                  # neighbour draw stays on the GPU (no host sync in the step)
    opt.render.device_rng = True                    # stratified jitter / eikonal samples from the CUDA generator
    r = measure_train_step(a, opt, a.batch, a.steps, a.warmup, rank, world, dev)
    if rank != 0:
        scdist.cleanup()
        return
    pk = peaks()
    line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=world, steps=a.steps, warmup=max(3, a.warmup),
                ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(a.batch, a.gpus, opt, clip=r["clip"]), impl="ours", e2e=r["e2e"],
                gpu_launches=r["gpu_launches"], clocks=r["clocks"], roofline=r["roofline"])
    line["config"]["execution"] = "eager launches" if a.eager else "ONE CUDA graph per step (forward + losses + backward + all-reduce + Adam) replayed"
    if r["allreduce"] is not None:
        line["allreduce"] = r["allreduce"]
    if world == 1 and not a.no_cpu_baseline:
        rate, dt, cores, kind = cpu_reference_rate(1, 1, 1)
        line["cpu_baseline"] = dict(value=rate, unit=UNIT, cores=cores, kind=kind,
                                    sample="1 image/step of the same workload (CLIP ViT-B/32 encode + top-6, 512 rays x 64 samples, 2 renders, "
                                           "losses, backward, Adam), 1 warm-up + 1 timed step; " +
                                           ("the reference's own Python modules (oracle/_ref/py)" if kind == "reference" else "oracle/ port"))
    if world == 1 and not a.no_configs:
        cfgs = {}
        for name, fn in (("configs[0]", config0_record),
                         ("configs[1]", lambda: _config1(a, opt, dev, r)),
                         ("configs[2]", lambda: config2_record(a, dev, pk)),
                         ("configs[4]", lambda: config4_record(a, dev, pk)),
                         ("ray_sampler", lambda: ray_sampler_record(a, dev, pk)),
                         ("step_with_clip_fp16", lambda: _step_clip_fp16(a, opt, dev))):
            try:
                cfgs[name] = fn()
            except Exception as ex:  # noqa: BLE001 - a sub-record never takes the headline line down with it
                import traceback
                cfgs[name] = dict(error=repr(ex), trace=traceback.format_exc()[-1500:])
        cfgs["configs[3]"] = "the headline of this line at --gpus 8 (32 images per GPU = 256 global); see SCALE records"
        line["configs"] = cfgs
    print(json.dumps(line), flush=True)
    scdist.cleanup()


def _config1(a, opt, dev, main):
    if a.batch == 16:
        return dict(workload="configs[1]: this line's headline")
    r = measure_train_step(a, opt, 16, steps=min(a.steps, 20), warmup=3, rank=0, world=1, dev=dev, sample_clocks=False)
    return dict(workload=workload_config(16, 1, opt, clip=r["clip"])["workload"], images_per_s=r["value"], ms_per_step=r["ms_per_step"],
                e2e=r["e2e"], gpu_launches=r["gpu_launches"], roofline=r["roofline"])


def _step_clip_fp16(a, opt, dev):
    """The headline step with the CLIP leg in the tower's fp16-operand mode: the arithmetic the reference itself runs CLIP at on CUDA
    (`clip.load` returns an fp16 model, CLIP_anno.py:16; here with fp32 accumulation and an fp32 residual stream, 2.7e-4 on the
    embedding). The headline keeps the fp32-class parity mode (hi/lo bf16 pairs, 3 MMAs per product, 1e-4)."""
    old = os.environ.get("SC_BENCH_CLIP")
    os.environ["SC_BENCH_CLIP"] = "fp16"
    try:
        r = measure_train_step(a, opt, a.batch, steps=min(a.steps, 20), warmup=3, rank=0, world=1, dev=dev, sample_clocks=False)
    finally:
        if old is None:
            os.environ.pop("SC_BENCH_CLIP", None)
        else:
            os.environ["SC_BENCH_CLIP"] = old
    return dict(workload=workload_config(a.batch, 1, opt, clip=r["clip"])["workload"] + "; CLIP leg in fp16 operands (the reference's own CLIP precision)",
                images_per_s=r["value"], ms_per_step=r["ms_per_step"], e2e_images_per_s=r["e2e"]["value"])


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
