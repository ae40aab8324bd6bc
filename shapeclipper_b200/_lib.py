"""ctypes loader for libsc_b200.so (the C ABI in include/sc_b200.h). There is no CPU fallback: if the
library is missing or a call fails, the caller gets an exception."""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libsc_b200.so")
ABI_VERSION = 4
_lib = None

c_float_p = ctypes.c_void_p   # raw device addresses travel as void*
c_int = ctypes.c_int
c_size_t = ctypes.c_size_t
c_void_p = ctypes.c_void_p


class NativeLibraryError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise NativeLibraryError(
                "%s not found: build it with `python -m shapeclipper_b200.build` "
                "(shapeclipper_b200 has no CPU or PyTorch fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.sc_abi_version.restype = c_int
        if L.sc_abi_version() != ABI_VERSION:
            raise NativeLibraryError("libsc_b200.so ABI %d != expected %d: rebuild" % (L.sc_abi_version(), ABI_VERSION))
        _declare(L)
        _lib = L
    return _lib


def _declare(L):
    vp, i, sz = c_void_p, c_int, c_size_t
    L.sc_chamfer_workspace_bytes.argtypes = [i, i, i]
    L.sc_chamfer_workspace_bytes.restype = sz
    L.sc_chamfer_forward.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, vp, sz, vp]
    L.sc_chamfer_forward.restype = i
    L.sc_chamfer_backward.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.sc_chamfer_backward.restype = i
    L.sc_pixel_rays_forward.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp]
    L.sc_pixel_rays_forward.restype = i
    L.sc_pixel_rays_backward.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    L.sc_pixel_rays_backward.restype = i
    f, d = ctypes.c_float, ctypes.c_double
    L.sc_eikonal_points_forward.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, f, f, vp, vp, vp]
    L.sc_eikonal_points_forward.restype = i
    L.sc_eikonal_points_backward.argtypes = [vp, vp, vp, i, i, f, vp, vp, vp]
    L.sc_eikonal_points_backward.restype = i
    L.sc_render_losses_workspace_floats.argtypes = [i]
    L.sc_render_losses_workspace_floats.restype = sz
    L.sc_render_losses_pass1.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, f, vp, vp, vp, vp, vp, vp, vp, vp]
    L.sc_render_losses_pass1.restype = i
    L.sc_render_losses_pass2.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, f, d, f, vp, vp, vp, vp, vp, vp]
    L.sc_render_losses_pass2.restype = i
    from . import _render_native, clip, mcubes, sampling
    _render_native.declare(L)
    clip.declare(L)
    mcubes.declare(L)
    sampling.declare(L)


def ptr(t):
    """Device address of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def check(code, what):
    if code != 0:
        raise NativeLibraryError("%s failed with cudaError %d" % (what, code))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise NativeLibraryError("shapeclipper_b200 runs on CUDA tensors only (got a %s tensor)" % t.device)
