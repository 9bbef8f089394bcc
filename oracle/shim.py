"""Compatibility shim that imports the UNMODIFIED lang2seg reference modules.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lang2seg_b200/`` may import this
file.  It is used (i) by ``oracle/make_golden.py`` in the development
container to generate the fixtures under ``tests/golden/`` and (ii) by the
``-m "not gpu"`` tests that pin ``oracle/restate.py`` against the real
reference when ``/root/reference`` is mounted.  The reference tree does not
exist on the GPU box; callers must check :func:`available` first.

Every item below is a recorded deviation:

D1  sys.path += lib/, pyutils/mask-faster-rcnn/lib/   (tools/_init_paths.py:5-16)
D2  stub modules easydict, tensorboardX, pycocotools.mask, _ext.{roi_pooling,nms}
D3  F.affine_grid / F.grid_sample forced to align_corners=True (PyTorch 0.3 semantics,
    pyutils/mask-faster-rcnn/lib/nets/network_cycle_response.py:142-147)
D4  scipy.misc.imresize -> PIL resize (uint8 is not rescaled)
D5  np.float = float ; Tensor.cuda / Module.cuda -> identity on the CPU run
D6  entry at Network._predict and sub-methods, not forward/train_step
D7  _region_proposal replaced by caller-supplied ROIs ; _anchor_component no-op
D8  .eval() for every parity run (dropout off)
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

def _default_root():
    """/root/reference where it is mounted (development container); else the copy that oracle/install_reference.py left
    under baseline/_ref (git-ignored, travels to the GPU box)."""
    if os.path.isdir("/root/reference/lib"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


REF_ROOT = os.environ.get("L2S_REFERENCE_ROOT") or _default_root()
_MFR = os.path.join(REF_ROOT, "pyutils", "mask-faster-rcnn", "lib")
_LIB = os.path.join(REF_ROOT, "lib")

_installed = False


def available() -> bool:
    return os.path.isdir(_MFR) and os.path.isdir(_LIB)


class _AttrDict(dict):
    """easydict stand-in (D2): nested dicts become attribute dicts."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _AttrDict):
            v = _AttrDict(v)
        dict.__setitem__(self, k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e


def _imresize(arr, size, interp="bilinear", mode=None):
    """scipy.misc.imresize replacement (D4)."""
    from PIL import Image

    arr = np.asarray(arr)
    if arr.dtype != np.uint8:
        # scipy's bytescale for non-uint8 input
        lo, hi = float(arr.min()), float(arr.max())
        scale = 255.0 / (hi - lo) if hi > lo else 1.0
        arr = ((arr - lo) * scale + 0.5).clip(0, 255).astype(np.uint8)
    if isinstance(size, (int, np.integer)):
        h, w = [int(s * size / 100.0) for s in arr.shape[:2]]
    elif isinstance(size, float):
        h, w = [int(s * size) for s in arr.shape[:2]]
    else:
        h, w = int(size[0]), int(size[1])
    flt = {"nearest": Image.NEAREST, "bilinear": Image.BILINEAR, "bicubic": Image.BICUBIC}[interp]
    return np.asarray(Image.fromarray(arr).resize((w, h), flt))


import contextlib


@contextlib.contextmanager
def cpu_only():
    """D5, scoped: `Tensor.cuda()` / `Module.cuda()` are the identity while the reference modules run as the CPU baseline
    on a box that HAS a GPU (install() only patches them for good where CUDA is absent); restored on exit so that the
    caller's own GPU code is untouched."""
    t, m = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = t, m


def install():
    """Install the stubs / patches once and extend sys.path."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)

    for p in (_MFR, _LIB):
        if p not in sys.path:
            sys.path.insert(0, p)

    # D2 ---------------------------------------------------------------
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _AttrDict
        sys.modules["easydict"] = m
    if "tensorboardX" not in sys.modules:
        m = types.ModuleType("tensorboardX")
        m.summary = types.SimpleNamespace(
            image=lambda *a, **k: None, histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)
        m.FileWriter = object
        sys.modules["tensorboardX"] = m
    if "pycocotools" not in sys.modules:
        m = types.ModuleType("pycocotools")
        mm = types.ModuleType("pycocotools.mask")
        m.mask = mm
        sys.modules["pycocotools"] = m
        sys.modules["pycocotools.mask"] = mm
    if "_ext" not in sys.modules:
        m = types.ModuleType("_ext")
        m.roi_pooling = types.ModuleType("_ext.roi_pooling")
        m.nms = types.ModuleType("_ext.nms")
        sys.modules["_ext"] = m
        sys.modules["_ext.roi_pooling"] = m.roi_pooling
        sys.modules["_ext.nms"] = m.nms
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except Exception:
            sys.modules["cv2"] = types.ModuleType("cv2")

    # D4 ---------------------------------------------------------------
    import scipy.misc

    scipy.misc.imresize = _imresize

    # D5 ---------------------------------------------------------------
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    # D3 ---------------------------------------------------------------
    _ag, _gs = F.affine_grid, F.grid_sample

    def affine_grid(theta, size, align_corners=None):
        return _ag(theta, size, align_corners=True)

    def grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=None):
        return _gs(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=True)

    F.affine_grid = affine_grid
    F.grid_sample = grid_sample
    _installed = True


DEFAULT_OPT = dict(
    # tools/opt_cycle.py defaults + refcoco_unc/caption_log_response/infos-best.pkl
    vocab_size=1999, word_embedding_size=512, word_vec_size=512, rnn_hidden_size=512,
    bidirectional=1, word_drop_out=0.5, rnn_drop_out=0.2, rnn_num_layers=1, rnn_type="lstm",
    variable_lengths=1, C4_feat_dim=1024, cap_loss_weight=1.0,
    caption_model="att2in2", input_encoding_size=512, rnn_size=512, num_layers=1,
    drop_prob_lm=0.5, seq_length=10, fc_feat_size=4096, att_feat_size=4096, att_hid_size=512,
    start_from=None, dataset_splitBy="refcoco_unc",
)


def build_reference_net(opt=None, num_layers=101, seed=0):
    """Instantiate the reference resnetv1 (cycle+response) net on CPU, eval mode (D8)."""
    install()
    from nets.resnet_v1_cycle_response import resnetv1

    o = dict(DEFAULT_OPT)
    o.update(opt or {})
    torch.manual_seed(seed)
    net = resnetv1(o, batch_size=1, num_layers=num_layers)
    net.create_architecture(81, tag="default", anchor_scales=(4, 8, 16, 32), anchor_ratios=(0.5, 1, 2))
    net.eval()
    return net


def reference_caption_model(opt=None, seed=0):
    install()
    import caption_models

    o = dict(DEFAULT_OPT)
    o.update(opt or {})
    torch.manual_seed(seed)
    m = caption_models.setup(o)
    m.eval()
    return m


def reference_rnn_encoder(opt=None, seed=0):
    install()
    from layers.lang_encoder import RNNEncoder

    o = dict(DEFAULT_OPT)
    o.update(opt or {})
    torch.manual_seed(seed)
    enc = RNNEncoder(vocab_size=o["vocab_size"], word_embedding_size=o["word_embedding_size"],
                     word_vec_size=o["word_vec_size"], hidden_size=o["rnn_hidden_size"],
                     bidirectional=o["bidirectional"] > 0, input_dropout_p=o["word_drop_out"],
                     dropout_p=o["rnn_drop_out"], n_layers=o["rnn_num_layers"], rnn_type=o["rnn_type"],
                     variable_lengths=o["variable_lengths"] > 0)
    enc.eval()
    return enc


def run_predict(net, net_conv, labels, rois, mode="TEST", num_fg=None):
    """Run the reference ``Network._predict`` with the backbone and RPN stubbed (D6, D7).

    net_conv : (1,C,H,W) float tensor standing in for ``_image_to_head()``
    labels   : (1,L) int64 expression tokens
    rois     : (R,5) float [0,x1,y1,x2,y2]
    """
    from torch.autograd import Variable  # noqa: F401

    net._mode = mode
    net._labels = labels
    net._image_to_head = lambda: net_conv
    net._anchor_component = lambda h, w: None
    net._region_proposal = lambda nc: rois
    net._im_info = np.array([[net_conv.shape[2] * 16, net_conv.shape[3] * 16, 1.0]], dtype=np.float32)
    if mode == "TRAIN":
        net._proposal_targets["mask_targets"] = torch.zeros(num_fg, 14, 14)
    return net._predict()
