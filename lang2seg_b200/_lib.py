"""ctypes binding of libl2s.so (include/l2s.h).  There is NO fallback: if the library is
missing, cannot be loaded, or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libl2s.so")

_vp, _i, _f, _sz, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_int64

# name -> (restype, argtypes) ; mirrors include/l2s.h exactly (tests/test_abi.py checks the symbol list)
SIGNATURES = {
    "l2s_version": (_i, []),
    "l2s_last_error_string": (ctypes.c_char_p, []),
    "l2s_launch_count": (ctypes.c_uint64, []),
    "l2s_set_debug_buffer": (_i, [_vp, _sz]),
    "l2s_set_precision": (_i, [_i]),
    "l2s_get_precision": (_i, []),
    "l2s_dynfilter_fwd_workspace_bytes": (_sz, [_i] * 5),
    "l2s_dynfilter_fwd": (_i, [_vp] * 9 + [_i] * 6 + [_vp, _sz, _vp]),
    "l2s_dynfilter_bwd_workspace_bytes": (_sz, [_i] * 5),
    "l2s_dynfilter_bwd": (_i, [_vp] * 13 + [_i] * 6 + [_vp, _sz, _vp]),
    "l2s_roi_crop_workspace_bytes": (_sz, [_i, _i, _i]),
    "l2s_roi_crop_fwd": (_i, [_vp] * 4 + [_i] * 7 + [_f, _f, _vp, _sz, _vp]),
    "l2s_roi_crop_bwd": (_i, [_vp] * 4 + [_i] * 7 + [_f, _f, _vp, _sz, _vp]),
    "l2s_roi_maxpool_fwd": (_i, [_i, _i, _f] + [_vp] * 4 + [_i] * 5 + [_vp]),
    "l2s_roi_maxpool_bwd": (_i, [_i, _i, _f] + [_vp] * 4 + [_i] * 5 + [_vp]),
    "l2s_split_bf16": (_i, [_vp] * 3 + [_i64] * 4 + [_vp]),
    "l2s_gemm_bf16x3": (_i, [_vp] * 6 + [_i] * 8 + [_vp]),
    "l2s_gemm_f32": (_i, [_vp] * 3 + [_i] * 3 + [_i64] * 5 + [_i, _vp]),
    "l2s_mask_head_saved_bytes": (_sz, [_i] * 4),
    "l2s_mask_head_workspace_bytes": (_sz, [_i] * 4),
    "l2s_mask_head_fwd": (_i, [_vp] * 8 + [_i] * 4 + [_vp, _sz, _vp]),
    "l2s_mask_head_fwd_stages": (_i, [_vp] * 8 + [_i] * 4 + [_vp, _sz, _i, _vp]),
    "l2s_mask_head_bwd": (_i, [_vp] * 9 + [_i] * 4 + [_vp, _sz, _vp]),
    "l2s_mask_head_bce_bwd": (_i, [_vp] * 12 + [_i] * 4 + [_vp, _sz, _vp]),
    "l2s_proposal_decode": (_i, [_vp, _vp, _vp, _i64, _vp, _i, _f, _f, _vp]),
    "l2s_roi_gt_overlaps": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "l2s_bbox_targets": (_i, [_vp, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "l2s_spatial_mean_fwd": (_i, [_vp, _vp, _i64, _i, _vp]),
    "l2s_spatial_mean_bwd": (_i, [_vp, _vp, _i64, _i, _vp]),
    "l2s_softmax_argmax": (_i, [_vp, _i64, _vp, _vp, _i, _i, _vp]),
    "l2s_colsum_workspace_bytes": (_sz, [_i, _i]),
    "l2s_colsum": (_i, [_vp, _i64, _vp, _i, _i, _vp, _sz, _vp]),
    "l2s_nms_workspace_bytes": (_sz, [_i]),
    "l2s_nms": (_i, [_vp, _i, _f, _i, _vp, _vp, _vp, _sz, _vp]),
    "l2s_mask_crop_resize": (_i, [_vp, _vp, _i, _vp, _vp] + [_i] * 6 + [_vp]),
    "l2s_mask_bce_fwd": (_i, [_vp] * 4 + [_i] * 3 + [_vp]),
    "l2s_mask_bce_bwd": (_i, [_vp] * 5 + [_i] * 3 + [_vp]),
    "l2s_att_step_fwd": (_i, [_vp] * 7 + [_i] * 4 + [_vp]),
    "l2s_att_step_bwd": (_i, [_vp] * 11 + [_i] * 4 + [_vp]),
    "l2s_att2in2_gates_fwd": (_i, [_vp] * 5 + [_i] * 2 + [_vp]),
    "l2s_att2in2_gates_bwd": (_i, [_vp] * 9 + [_i] * 2 + [_vp]),
    "l2s_att2in2_decode_workspace_bytes": (_sz, [_i] * 5),
    "l2s_att2in2_decode_fwd": (_i, [_vp] * 13 + [_i] * 5 + [_vp, _sz, _vp]),
    "l2s_att2in2_decode_bwd": (_i, [_vp] * 17 + [_i] * 5 + [_vp, _sz, _vp]),
    "l2s_bilstm_workspace_bytes": (_sz, [_i] * 3),
    "l2s_bilstm_fwd": (_i, [_vp] * 7 + [_i] * 3 + [_vp, _sz, _vp]),
    "l2s_bilstm_bwd": (_i, [_vp] * 7 + [_i] * 3 + [_vp, _sz, _vp]),
    "l2s_embedding_bwd": (_i, [_vp] * 3 + [_i] * 3 + [_vp]),
    "l2s_linear_small_workspace_bytes": (_sz, [_i] * 3),
    "l2s_linear_small": (_i, [_vp] * 4 + [_i] * 7 + [_vp, _sz, _vp]),
    "l2s_logsoftmax_nll_fwd": (_i, [_vp] * 5 + [_i] * 2 + [_vp]),
    "l2s_logsoftmax_nll_bwd": (_i, [_vp] * 5 + [_i] * 2 + [_vp]),
    "l2s_caption_feats_fwd": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
    "l2s_caption_feats_bwd": (_i, [_vp] * 3 + [_i] * 7 + [_vp]),
}

_lib = None


def load():
    """Load libl2s.so; raises if it has not been built (python -m lang2seg_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "lang2seg_b200: %s is missing -- build it with `python -m lang2seg_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI drifted
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


class L2SError(RuntimeError):
    pass


def call(name, *args):
    """Invoke an int-returning entry point; raise L2SError(l2s_last_error_string()) on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise L2SError("%s failed (%d): %s" % (name, rc, lib.l2s_last_error_string().decode()))


def size(name, *args):
    return int(getattr(load(), name)(*args))


def launch_count():
    return int(load().l2s_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda, "lang2seg_b200 kernels need CUDA tensors (no CPU path)"
    return ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def f32c(t):
    """contiguous fp32 view/copy"""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
