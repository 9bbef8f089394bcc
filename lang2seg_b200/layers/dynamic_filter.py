"""Filter generator + spatial dynamic-filter response layer.

Drop-in for network_cycle_response.py:510-570 with the parameters of
resnet_v1_cycle_response.py:313-321 (dynamic_fc_0..6, response_fc) kept under the same names.
"""
import torch
import torch.nn as nn

from .. import functional as L2F


def generate_filters(hidden, dynamic_fcs, response_fc):
    """f_k = tanh(dynamic_fc_k(hidden)) stacked to (E,7,C) ; w = tanh(response_fc(hidden)) (E,7).  (:510-532)

    The reference issues 8 separate Linears per expression; here the seven (C x Dh) projections AND response_fc run as
    ONE GEMM over the stacked weights (7C + 7 rows, padded to a multiple of 8), forward and backward, on the library's
    kernels (L2F.dense: skinny exact-fp32 GEMM for small batches of expressions, tcgen05 above 256)."""
    E, Dh = hidden.shape
    nf = len(dynamic_fcs)
    C = dynamic_fcs[0].out_features
    nr = response_fc.out_features
    pad = (-(nf * C + nr)) % 8
    W = torch.cat([fc.weight for fc in dynamic_fcs] + [response_fc.weight] +
                  ([hidden.new_zeros(pad, Dh)] if pad else []), 0)
    b = torch.cat([fc.bias for fc in dynamic_fcs] + [response_fc.bias] + ([hidden.new_zeros(pad)] if pad else []), 0)
    y = torch.tanh(L2F.dense(hidden, W, b))
    return y[:, :nf * C].reshape(E, nf, C), y[:, nf * C:nf * C + nr]


class DynamicFilterResponse(nn.Module):
    """Owns dynamic_fc_0..6 / response_fc and applies the fused sm_100a response kernel."""

    def __init__(self, hidden_dim, feat_dim, gate="sigmoid"):
        super().__init__()
        for k in range(L2F.NUM_FILTERS):
            setattr(self, "dynamic_fc_%d" % k, nn.Linear(hidden_dim, feat_dim))
        self.response_fc = nn.Linear(hidden_dim, L2F.NUM_FILTERS)
        self.gate = gate

    @property
    def dynamic_fcs(self):
        return [getattr(self, "dynamic_fc_%d" % k) for k in range(L2F.NUM_FILTERS)]

    def forward(self, net_conv, hidden, expr2img=None, resp_target=None):
        filt, fuse = generate_filters(hidden, self.dynamic_fcs, self.response_fc)
        return L2F.dynamic_filter(net_conv, filt, fuse, expr2img, self.gate, resp_target)
