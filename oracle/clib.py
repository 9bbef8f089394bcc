"""ctypes access to the oracle's native pieces (TEST INFRASTRUCTURE ONLY).

liboracle_c.so        -- oracle/roi_ref.c (plain C restatement)
libroi_pooling_ref.so -- the reference's own RoI max-pool CUDA kernels compiled for sm_100a
                         (oracle/Makefile `ref`); GPU only, raw device pointers.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
C_LIB = os.path.join(HERE, "_build", "liboracle_c.so")
REF_LIB = os.path.join(HERE, "_ref", "libroi_pooling_ref.so")


def build(quiet=True):
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _c():
    if not os.path.exists(C_LIB):
        build()
    return ctypes.CDLL(C_LIB)


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def roi_maxpool_fwd(feat, rois, ph=7, pw=7, scale=1.0 / 16):
    feat = np.ascontiguousarray(feat, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, C, H, W = feat.shape
    N = rois.shape[0]
    out = np.empty((N, C, ph, pw), np.float32)
    arg = np.empty((N, C, ph, pw), np.int32)
    rc = _c().oracle_roi_maxpool_fwd(_fp(feat), _fp(rois), B, C, H, W, N, ph, pw, ctypes.c_float(scale),
                                     _fp(out), _ip(arg))
    assert rc == 0
    return out, arg


def roi_maxpool_bwd(top, rois, arg, shape):
    top = np.ascontiguousarray(top, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    arg = np.ascontiguousarray(arg, np.int32)
    B, C, H, W = shape
    N, _, ph, pw = top.shape
    out = np.empty(shape, np.float32)
    rc = _c().oracle_roi_maxpool_bwd(_fp(top), _fp(rois), _ip(arg), B, C, H, W, N, ph, pw, _fp(out))
    assert rc == 0
    return out


def crop_resize_fwd(feat, rois, S=7, max_pool=False, align_im_hw=None):
    feat = np.ascontiguousarray(feat, np.float32)
    rois = np.ascontiguousarray(rois, np.float32)
    B, C, H, W = feat.shape
    N = rois.shape[0]
    P = S // 2 if max_pool else S
    out = np.empty((N, C, P, P), np.float32)
    imh, imw = align_im_hw if align_im_hw is not None else (0.0, 0.0)
    rc = _c().oracle_crop_resize_fwd(_fp(feat), _fp(rois), B, C, H, W, N, S, int(max_pool),
                                     int(align_im_hw is not None), ctypes.c_float(imh), ctypes.c_float(imw),
                                     _fp(out))
    assert rc == 0
    return out


def reference_roi_pool_cuda():
    """The reference's ROIPool{Forward,Backward}Laucher (roi_pooling_kernel.cu:78-101,181-202) or None."""
    if not os.path.exists(REF_LIB):
        return None
    lib = ctypes.CDLL(REF_LIB)
    vp, f, i = ctypes.c_void_p, ctypes.c_float, ctypes.c_int
    lib.ROIPoolForwardLaucher.argtypes = [vp, f, i, i, i, i, i, i, vp, vp, vp, vp]
    lib.ROIPoolBackwardLaucher.argtypes = [vp, f, i, i, i, i, i, i, i, vp, vp, vp, vp]
    return lib


REF_NMS_LIB = os.path.join(HERE, "_ref", "libnms_ref.so")


def reference_nms_cuda():
    """The reference's `_nms(boxes_num, boxes_dev, mask_dev, thresh)` launcher (nms/src/cuda/nms_kernel.cu:72-83,
    default stream) or None when oracle/_ref/libnms_ref.so has not been built."""
    if not os.path.exists(REF_NMS_LIB):
        return None
    lib = ctypes.CDLL(REF_NMS_LIB)
    lib._nms.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]
    lib._nms.restype = None
    return lib
