"""ctypes mirror of ScRenderArgs (include/sc_b200.h) and thin launch helpers. Device pointers only; every
launch goes to the current torch stream of the tensors' device. No fallback: errors raise."""
import ctypes

import torch

from . import _lib

_fp = ctypes.c_void_p


class ScRenderArgs(ctypes.Structure):
    _fields_ = [
        ("mode", ctypes.c_int), ("batch", ctypes.c_int), ("n_per_image", ctypes.c_int), ("n_samples", ctypes.c_int),
        ("want_grad", ctypes.c_int), ("want_feat", ctypes.c_int), ("detach_latent", ctypes.c_int),
        ("beta_min", ctypes.c_float), ("cam_dist", ctypes.c_float), ("half_range", ctypes.c_float),
        ("bg_color", ctypes.c_float), ("normal_pow", ctypes.c_float),
        ("blob", _fp), ("cb", _fp), ("beta_param", _fp),
        ("cam_loc", _fp), ("ray_dirs", _fp), ("depth_fac", _fp), ("scale_dist", _fp), ("t_vals", _fp), ("jitter", _fp),
        ("points", _fp),
        ("rgb", _fp), ("mask", _fp), ("mask_hard", _fp), ("depth", _fp), ("normal", _fp),
        ("sdf", _fp), ("feat", _fp), ("grad", _fp),
        ("rgb_bar", _fp), ("mask_bar", _fp), ("depth_bar", _fp), ("normal_bar", _fp), ("sdf_bar", _fp), ("grad_bar", _fp),
        ("grad_partial", _fp), ("cb_bar", _fp), ("ray_dirs_bar", _fp), ("depth_fac_bar", _fp), ("cam_loc_bar", _fp),
        ("scale_dist_bar", _fp), ("points_bar", _fp),
        ("scratch", _fp), ("saved", _fp), ("precision", ctypes.c_int),
    ]


def declare(L):
    vp, i, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    L.sc_render_blob_floats.restype = sz
    L.sc_render_grad_floats.restype = sz
    L.sc_render_num_ctas.restype = i
    L.sc_render_scratch_bytes.argtypes = [i]
    L.sc_render_scratch_bytes.restype = sz
    L.sc_render_pack_weights.argtypes = [vp, vp, vp, vp]
    L.sc_render_pack_weights.restype = i
    L.sc_render_latent_bias.argtypes = [vp, vp, vp, i, vp, vp]
    L.sc_render_latent_bias.restype = i
    L.sc_render_forward.argtypes = [ctypes.POINTER(ScRenderArgs), vp]
    L.sc_render_forward.restype = i
    L.sc_render_tc_blob_bytes.restype = sz
    L.sc_render_tc_scratch_bytes.argtypes = [i]
    L.sc_render_tc_scratch_bytes.restype = sz
    L.sc_render_tc_pack_weights.argtypes = [vp, vp, vp, vp, vp]
    L.sc_render_tc_pack_weights.restype = i
    L.sc_render_tc_saved_bytes.argtypes = [i, i, i]
    L.sc_render_tc_saved_bytes.restype = sz
    L.sc_render_tc_forward.argtypes = [ctypes.POINTER(ScRenderArgs), vp]
    L.sc_render_tc_forward.restype = i
    L.sc_render_tc2_supported.argtypes = [i, i]
    L.sc_render_tc2_supported.restype = i
    L.sc_render_tc2_forward.argtypes = [ctypes.POINTER(ScRenderArgs), vp]
    L.sc_render_tc2_forward.restype = i
    L.sc_render_tc2_backward.argtypes = [ctypes.POINTER(ScRenderArgs), vp]
    L.sc_render_tc2_backward.restype = i
    if hasattr(L, "sc_render_tc_backward"):
        L.sc_render_tc_backward.argtypes = [ctypes.POINTER(ScRenderArgs), vp]
        L.sc_render_tc_backward.restype = i
    if hasattr(L, "sc_render_backward"):
        L.sc_render_backward.argtypes = [ctypes.POINTER(ScRenderArgs), vp]
        L.sc_render_backward.restype = i
        L.sc_render_grad_finalize.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp]
        L.sc_render_grad_finalize.restype = i
        L.sc_render_grad_finalize_accumulate.argtypes = L.sc_render_grad_finalize.argtypes
        L.sc_render_grad_finalize_accumulate.restype = i


def _p(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError("shapeclipper_b200 kernels take float32 tensors (got %s)" % t.dtype)
    return t.contiguous()


def pack_weights(weights, biases):
    """weights/biases: 10 CUDA fp32 tensors each (sdf lin0..5, rgb lin0..3, nn.Linear layout) -> blob tensor."""
    L = _lib.lib()
    dev = weights[0].device
    _lib.require_cuda(*weights, *biases)
    ws = [_f32c(w.detach()) for w in weights]
    bs = [_f32c(b.detach()) for b in biases]
    blob = torch.empty(L.sc_render_blob_floats(), dtype=torch.float32, device=dev)
    warr = (ctypes.c_void_p * 10)(*[w.data_ptr() for w in ws])
    barr = (ctypes.c_void_p * 10)(*[b.data_ptr() for b in bs])
    with torch.cuda.device(dev):
        _lib.check(L.sc_render_pack_weights(warr, barr, _p(blob), _lib.stream_of(blob)), "sc_render_pack_weights")
    TIMERS.count()
    return blob


_blob_cache = {}
_uid_counter = [0]


def _uid(t):
    """Identity of a parameter object that is never recycled (data_ptr / id() are, by the caching allocator / GC)."""
    u = getattr(t, "_sc_uid", None)
    if u is None:
        _uid_counter[0] += 1
        u = _uid_counter[0]
        try:
            t._sc_uid = u
        except Exception:  # noqa: BLE001 - tensors that cannot carry attributes are simply never cached
            return -_uid_counter[0]
    return u


def packed_blob(weights, biases):
    """pack_weights with a small cache keyed on (storage, version) of every tensor: the kernel calls of one
    training step share one packing per parameter set. In-place updates through autograd-visible ops (optimisers,
    load_state_dict) bump the version; after raw `.data` writes call invalidate_blob_cache()."""
    key = tuple((_uid(t), t.data_ptr(), t._version) for t in list(weights) + list(biases))
    dev = str(weights[0].device)
    cache = _blob_cache.setdefault(dev, {})
    blob = cache.get(key)
    if blob is None:
        if len(cache) >= 4:
            cache.clear()
        blob = pack_weights(weights, biases)
        cache[key] = blob
    return blob


def invalidate_blob_cache():
    _blob_cache.clear()
    _tc_blob_cache.clear()


_tc_blob_cache = {}


def packed_tc_blob(weights, biases, ffma_blob):
    """The tensor-core blob (swizzled hi/lo bf16 weight tiles) for the same parameter set, cached like packed_blob."""
    L = _lib.lib()
    key = tuple((_uid(t), t.data_ptr(), t._version) for t in list(weights) + list(biases))
    dev = weights[0].device
    cache = _tc_blob_cache.setdefault(str(dev), {})
    blob = cache.get(key)
    if blob is None:
        if len(cache) >= 4:
            cache.clear()
        ws = [_f32c(w.detach()) for w in weights]
        bs = [_f32c(b.detach()) for b in biases]
        blob = torch.empty(L.sc_render_tc_blob_bytes(), dtype=torch.uint8, device=dev)
        warr = (ctypes.c_void_p * 10)(*[w.data_ptr() for w in ws])
        barr = (ctypes.c_void_p * 10)(*[b.data_ptr() for b in bs])
        with torch.cuda.device(dev):
            _lib.check(L.sc_render_tc_pack_weights(warr, barr, _p(ffma_blob), _p(blob), _lib.stream_of(blob)),
                       "sc_render_tc_pack_weights")
        TIMERS.count()
        cache[key] = blob
    return blob


def latent_bias(blob, z_sdf, z_rgb, batch):
    L = _lib.lib()
    cb = torch.empty(batch, 4, 64, dtype=torch.float32, device=blob.device)
    zs = _f32c(z_sdf.detach()) if z_sdf is not None else None
    zr = _f32c(z_rgb.detach()) if z_rgb is not None else None
    with torch.cuda.device(blob.device):
        _lib.check(L.sc_render_latent_bias(_p(blob), _p(zs), _p(zr), batch, _p(cb), _lib.stream_of(blob)),
                   "sc_render_latent_bias")
    TIMERS.count()
    return cb


def scratch(device, backward, tc=False):
    L = _lib.lib()
    with torch.cuda.device(device):
        n = (L.sc_render_tc_scratch_bytes if tc else L.sc_render_scratch_bytes)(1 if backward else 0)
    return torch.empty(n // 4, dtype=torch.float32, device=device)


SAVE_ACTIVATIONS_MAX_BYTES = 24 << 30     # per render; above this the backward recomputes the forward per tile instead


def saved_buffer(device, batch, n_per_image, n_samples):
    """Activation buffer the generation-1 tensor-core forward fills for its backward (ScRenderArgs.saved), or None when the
    shape is not supported / too large (the backward then recomputes)."""
    L = _lib.lib()
    n = L.sc_render_tc_saved_bytes(int(batch), int(n_per_image), int(n_samples))
    if n == 0 or n > SAVE_ACTIVATIONS_MAX_BYTES:
        return None
    # several renders' buffers are alive at once (1 + n_views per step): leave half of what is free right now alone
    free, _ = torch.cuda.mem_get_info(device)
    reusable = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    if n > (free + reusable) // 2:
        return None
    try:
        return torch.empty(n // 4, dtype=torch.float32, device=device)
    except torch.cuda.OutOfMemoryError:
        return None                                    # the backward recomputes the forward per tile instead


def saved_chunk_images(device, batch, n_per_image, n_samples):
    """How many images' saved activations fit the budget of saved_buffer() at once (0: not even one; >= batch: all of them)."""
    per_image = _lib.lib().sc_render_tc_saved_bytes(1, int(n_per_image), int(n_samples))
    if per_image == 0:
        return 0
    free, _ = torch.cuda.mem_get_info(device)
    reusable = torch.cuda.memory_reserved(device) - torch.cuda.memory_allocated(device)
    cap = min(SAVE_ACTIVATIONS_MAX_BYTES, (free + reusable) // 2)
    return int(min(batch, cap // per_image))


def saved_buffer_bytes(batch, n_per_image, n_samples):
    """Bytes of saved activations one training render of this shape would keep for its backward (0 = the backward recomputes)."""
    n = _lib.lib().sc_render_tc_saved_bytes(int(batch), int(n_per_image), int(n_samples))
    return 0 if (n == 0 or n > SAVE_ACTIVATIONS_MAX_BYTES) else int(n)


class KernelTimers:
    """Optional CUDA-event timing of the main kernels on their launch stream (bench.py's roofline numbers).
    Also counts every launch of a kernel of this library (bench.py's gpu_launches)."""

    def __init__(self):
        self.enabled = False
        self.capturing = False      # inside a CUDA-graph capture (step.TrainStep): spans become external event nodes
        self.events = {}
        self.graph_events = {}
        self.launches = 0

    def reset(self):
        self.events = {}
        self.launches = 0

    def count(self, n=1):
        self.launches += n

    def span(self, name, device):
        return _Span(self, name, device)

    def totals_ms(self):
        torch.cuda.synchronize()
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.events.items()}

    def graph_ms(self):
        """Spans recorded while a CUDA graph was captured, as timed by the LAST replay: {name: (total ms, launches)}."""
        torch.cuda.synchronize()
        return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in self.graph_events.items()}


class _Span:
    def __init__(self, timers, name, device):
        self.t, self.name, self.device = timers, name, device

    def __enter__(self):
        if self.t.enabled:
            ext = self.t.capturing
            self.a = torch.cuda.Event(enable_timing=True, external=ext)
            self.b = torch.cuda.Event(enable_timing=True, external=ext)
            self.a.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if self.t.enabled:
            self.b.record(torch.cuda.current_stream(self.device))
            (self.t.graph_events if self.t.capturing else self.t.events).setdefault(self.name, []).append((self.a, self.b))
        return False


TIMERS = KernelTimers()


def _entry(L, direction, tc, args):
    """tc: False = FP32 FFMA kernels, True / "tc" = tensor-core generation 1, "tc2" = generation 2 (two 64-point chains per
    CTA; falls back to generation 1 for sample counts that do not fit a 64-point tile)."""
    if not tc:
        name = "sc_render_" + direction
    elif tc == "tc2" and L.sc_render_tc2_supported(args.mode, args.n_samples):
        name = "sc_render_tc2_" + direction
    else:
        name = "sc_render_tc_" + direction
    return getattr(L, name), name


def launch_forward(args, device, tc=False, span=None):
    L = _lib.lib()
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        with TIMERS.span(span or ("render_fwd" if args.mode == 0 else "sdf_query_fwd"), device):
            fn, name = _entry(L, "forward", tc, args)
            _lib.check(fn(ctypes.byref(args), stream), name)
    TIMERS.count()


def launch_backward(args, device, tc=False):
    L = _lib.lib()
    with torch.cuda.device(device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        with TIMERS.span("render_bwd" if args.mode == 0 else "sdf_query_bwd", device):
            fn, name = _entry(L, "backward", tc, args)
            _lib.check(fn(ctypes.byref(args), stream), name)
    TIMERS.count()
