"""lang2seg_b200 -- B200 (sm_100a) kernels for lang2seg's language-conditioned segmentation hot path.

The package mirrors the reference's module interfaces for that path only:

    lang2seg_b200.functional              autograd bindings of the C ABI (include/l2s.h)
    lang2seg_b200.layers.lang_encoder     RNNEncoder             (lib/layers/lang_encoder.py)
    lang2seg_b200.layers.dynamic_filter   filter generator + response layer
    lang2seg_b200.layers.roi_pooling      RoIPoolFunction / RoIPool (layer_utils/roi_pooling/roi_pool.py)
    lang2seg_b200.caption_models          setup(opt) -> Att2in2Model (lib/caption_models)
    lang2seg_b200.misc.utils              LanguageModelCriterion  (lib/misc/utils.py)
    lang2seg_b200.nets.network            Network hot-path methods (_crop_pool_layer, _mask_prediction, ...)
    lang2seg_b200.parallel                data-parallel gradient all-reduce (NCCL)

Everything computes through libl2s.so; there is no CPU path and no PyTorch fallback.
"""
import torch as _torch

# fp32 parity (1e-4) on this path: keep cuDNN (the LSTM of the expression encoder) and cuBLAS off TF32.
_torch.backends.cudnn.allow_tf32 = False
_torch.backends.cuda.matmul.allow_tf32 = False

__version__ = "0.1.0"
