"""RoI max-pool -- drop-in for layer_utils/roi_pooling/roi_pool.py (RoIPoolFunction, RoIPool).

The reference's legacy instance-style Function (state on self, roi_pool.py:6-50) becomes a callable
object with the same constructor and call signature; the arithmetic runs in l2s_roi_maxpool_{fwd,bwd}.
"""
import torch.nn as nn

from .. import functional as L2F


class RoIPoolFunction(object):
    def __init__(self, pooled_height, pooled_width, spatial_scale):
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)
        self.argmax = None

    def __call__(self, features, rois):
        out, self.argmax = L2F.roi_max_pool(features, rois, self.pooled_height, self.pooled_width,
                                            self.spatial_scale, return_argmax=True)
        return out


class RoIPool(nn.Module):
    def __init__(self, pooled_height, pooled_width, spatial_scale):
        super().__init__()
        self.pooled_width = int(pooled_width)
        self.pooled_height = int(pooled_height)
        self.spatial_scale = float(spatial_scale)

    def forward(self, features, rois):
        return RoIPoolFunction(self.pooled_height, self.pooled_width, self.spatial_scale)(features, rois)
