"""Drop-in for the reference's pybind11 module `chamfer_3D` (external/chamfer3D/chamfer_cuda.cpp:30-33).

    forward(xyz1, xyz2, dist1, dist2, idx1, idx2) -> int      (1 ok / 0 error, like the original)
    backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2) -> int

Caller allocates all outputs (utils/eval_3D.py:155-165); results are written in place. Unlike the
original this launches on the CURRENT torch stream of the inputs' device rather than the legacy stream.
"""
import torch

from . import _lib


def _prep(xyz1, xyz2):
    _lib.require_cuda(xyz1, xyz2)
    if xyz1.dtype != torch.float32 or xyz2.dtype != torch.float32:
        raise TypeError("chamfer_3D expects float32 point clouds")
    if xyz1.dim() != 3 or xyz2.dim() != 3 or xyz1.shape[2] != 3 or xyz2.shape[2] != 3 or xyz1.shape[0] != xyz2.shape[0]:
        raise ValueError("expected xyz1 [B,N,3] and xyz2 [B,M,3]")
    return xyz1.contiguous(), xyz2.contiguous()


def forward(xyz1, xyz2, dist1, dist2, idx1, idx2):
    a, b = _prep(xyz1, xyz2)
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    for t, shape, dt in ((dist1, (B, N), torch.float32), (dist2, (B, M), torch.float32),
                         (idx1, (B, N), torch.int32), (idx2, (B, M), torch.int32)):
        if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.device != a.device:
            raise ValueError("output tensor must be contiguous %s %s on %s" % (shape, dt, a.device))
    L = _lib.lib()
    with torch.cuda.device(a.device):
        nbytes = L.sc_chamfer_workspace_bytes(B, N, M)
        ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=a.device)
        code = L.sc_chamfer_forward(_lib.ptr(a), _lib.ptr(b), B, N, M, _lib.ptr(dist1), _lib.ptr(dist2),
                                    _lib.ptr(idx1), _lib.ptr(idx2), _lib.ptr(ws), nbytes, _lib.stream_of(a))
    if code != 0:
        print("error in nnd updateOutput: cudaError %d" % code)
        return 0
    return 1


def backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
    a, b = _prep(xyz1, xyz2)
    B, N, M = a.shape[0], a.shape[1], b.shape[1]
    L = _lib.lib()
    with torch.cuda.device(a.device):
        code = L.sc_chamfer_backward(_lib.ptr(a), _lib.ptr(b), B, N, M,
                                     _lib.ptr(graddist1.contiguous()), _lib.ptr(graddist2.contiguous()),
                                     _lib.ptr(idx1.contiguous()), _lib.ptr(idx2.contiguous()),
                                     _lib.ptr(gradxyz1), _lib.ptr(gradxyz2), _lib.stream_of(a))
    if code != 0:
        print("error in nnd get grad: cudaError %d" % code)
        return 0
    return 1
