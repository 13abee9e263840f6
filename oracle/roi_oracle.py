"""oracle/roi_oracle.py -- ctypes front end of the ROI checkers.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).

  * liboracle.so        : oracle/roi_oracle.c, the CPU restatement (numpy arrays in / out)
  * _ref/libref_roi.so  : the reference's own vendored CUDA launchers compiled for sm_100a
                          (ROIAlignForwardLaucher / ROIAlignBackwardLaucher,
                          lib/modeling/roi_xfrom/roi_align/src/roi_align_kernel.cu:123-141,272-290;
                          ROIPoolForwardLaucher / ROIPoolBackwardLaucher,
                          lib/model/roi_pooling/src/roi_pooling_kernel.cu:96-125,206-239);
                          torch CUDA tensors in / out.  Semantics: aligned=False.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_cpu = None
_ref = None


def build(quiet=True):
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def cpu_lib():
    global _cpu
    if _cpu is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", HERE, "-s", "liboracle.so"], check=True)
        _cpu = C.CDLL(path)
    return _cpu


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def roi_align_fwd(feat, rois, oh, ow, scale, sampling_ratio=0, aligned=True):
    feat, rois = _f(feat), _f(rois)
    B, Cc, H, W = feat.shape
    K = rois.shape[0]
    out = np.empty((K, Cc, oh, ow), np.float32)
    cpu_lib().oracle_roi_align_fwd(_p(feat), _p(rois), _p(out), B, Cc, H, W, K, oh, ow, C.c_float(scale),
                                   int(sampling_ratio), int(bool(aligned)))
    return out


def roi_align_bwd(grad_out, rois, shape, scale, sampling_ratio=0, aligned=True):
    grad_out, rois = _f(grad_out), _f(rois)
    B, Cc, H, W = shape
    K, _, oh, ow = grad_out.shape
    gf = np.empty((B, Cc, H, W), np.float32)
    cpu_lib().oracle_roi_align_bwd(_p(grad_out), _p(rois), _p(gf), B, Cc, H, W, K, oh, ow, C.c_float(scale),
                                   int(sampling_ratio), int(bool(aligned)))
    return gf


def roi_pool_fwd(feat, rois, oh, ow, scale, variant="legacy"):
    """variant "legacy": the reference's vendored kernel (== torchvision.ops.roi_pool); "mmcv": mmcv 1.x's bins
    (restated from its published kernel; parity unpinned, see roi_oracle.c)."""
    feat, rois = _f(feat), _f(rois)
    B, Cc, H, W = feat.shape
    K = rois.shape[0]
    out = np.empty((K, Cc, oh, ow), np.float32)
    arg = np.empty((K, Cc, oh, ow), np.int32)
    fn = cpu_lib().oracle_roi_pool_fwd_mmcv if variant == "mmcv" else cpu_lib().oracle_roi_pool_fwd
    fn(_p(feat), _p(rois), _p(out), _p(arg), B, Cc, H, W, K, oh, ow, C.c_float(scale))
    return out, arg


def roi_pool_bwd(grad_out, argmax, rois, shape):
    grad_out, rois = _f(grad_out), _f(rois)
    argmax = np.ascontiguousarray(argmax, dtype=np.int32)
    B, Cc, H, W = shape
    K, _, oh, ow = grad_out.shape
    gf = np.empty((B, Cc, H, W), np.float32)
    cpu_lib().oracle_roi_pool_bwd(_p(grad_out), _p(argmax), _p(rois), _p(gf), B, Cc, H, W, K, oh, ow)
    return gf


def mask_counts(masks_u8):
    m = np.ascontiguousarray(masks_u8, dtype=np.uint8).reshape(len(masks_u8), -1)
    n, hw = m.shape
    inter = np.empty((n, n), np.int32)
    area = np.empty((n,), np.int32)
    cpu_lib().oracle_mask_counts(_p(m), n, C.c_int64(hw), _p(inter), _p(area))
    return inter, area


# ---------------------------------------------------------------- the reference's own kernels
def ref_lib():
    """oracle/_ref/libref_roi.so or None when it was never built (no reference tree, no nvcc)."""
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libref_roi.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
    return _ref


def _tp(t):
    return C.c_void_p(t.data_ptr())


def ref_roi_align_fwd(feat, rois, oh, ow, scale, sampling_ratio=0):
    """Vendored ROIAlignForward on the GPU (aligned=False semantics)."""
    import torch
    lib = ref_lib()
    B, Cc, H, W = feat.shape
    K = rois.shape[0]
    out = torch.empty((K, Cc, oh, ow), dtype=torch.float32, device=feat.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.ROIAlignForwardLaucher(_tp(feat), C.c_float(scale), K, H, W, Cc, oh, ow, int(sampling_ratio), _tp(rois),
                               _tp(out), st)
    return out


def ref_roi_align_bwd(grad_out, rois, shape, scale, sampling_ratio=0):
    import torch
    lib = ref_lib()
    B, Cc, H, W = shape
    K, _, oh, ow = grad_out.shape
    gf = torch.zeros((B, Cc, H, W), dtype=torch.float32, device=grad_out.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.ROIAlignBackwardLaucher(_tp(grad_out), C.c_float(scale), B, K, H, W, Cc, oh, ow, int(sampling_ratio),
                                _tp(rois), _tp(gf), st)
    return gf


def ref_roi_pool_fwd(feat, rois, oh, ow, scale):
    import torch
    lib = ref_lib()
    B, Cc, H, W = feat.shape
    K = rois.shape[0]
    out = torch.empty((K, Cc, oh, ow), dtype=torch.float32, device=feat.device)
    arg = torch.empty((K, Cc, oh, ow), dtype=torch.int32, device=feat.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.ROIPoolForwardLaucher(_tp(feat), C.c_float(scale), K, H, W, Cc, oh, ow, _tp(rois), _tp(out), _tp(arg), st)
    return out, arg
