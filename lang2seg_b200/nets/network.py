"""Hot-path half of the reference `Network` (pyutils/mask-faster-rcnn/lib/nets/network_cycle_response.py).

Only the methods on the path of SURVEY.md section 8 are provided, under the reference's names and
argument orders, so that a reference Network subclass can inherit them:

    _crop_pool_layer(bottom, rois, max_pool=True)                  :107-149
    _crop_pool_layer_align(bottom, rois, im_info, max_pool=True)   :151-182
    _roi_pool_layer(bottom, rois)                                  :104-105
    _region_classification(spatial_fc7)                            :277-290
    _mask_prediction(spatial_fc7)                                  :292-307
    _dynamic_filter(net_conv, labels | hidden)                     :503-570
    _add_hot_path_losses(...)                                      :375-453 (cls, box, mask, response, caption; weights :449)

`HotPathNet` is the concrete module (parameter names of resnet_v1_cycle_response.py:238-335) used by
bench.py and the tests; backbone, RPN and res5 are outside this repository's scope and are supplied
by the caller (`head_to_tail`) or replaced by synthetic tensors.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import caption_models
from .. import functional as L2F
from ..layers.dynamic_filter import generate_filters
from ..layers.lang_encoder import RNNEncoder
from ..layers.roi_pooling import RoIPoolFunction

POOLING_SIZE = 7      # cfg.POOLING_SIZE  (model/config.py:276)
MASK_SIZE = 14        # cfg.MASK_SIZE     (model/config.py:285)


class Network(nn.Module):
    def __init__(self, batch_size=1):
        super().__init__()
        self._feat_stride = [16, ]
        self._feat_compress = [1. / 16, ]
        self._batch_size = batch_size
        self._predictions = {}
        self._losses = {}
        self._proposal_targets = {}
        self._gate = "sigmoid"

    # ---- ROI pooling -------------------------------------------------------------------------
    def _roi_pool_layer(self, bottom, rois):
        return RoIPoolFunction(POOLING_SIZE, POOLING_SIZE, 1. / 16.)(bottom, rois)

    def _crop_pool_layer(self, bottom, rois, max_pool=True):
        return L2F.roi_crop(bottom, rois, max_pool=max_pool, pool=POOLING_SIZE)

    def _crop_pool_layer_align(self, bottom, rois, im_info, max_pool=True):
        im_h, im_w = float(im_info[0][0]), float(im_info[0][1])
        return L2F.roi_crop(bottom, rois, max_pool=max_pool, align_im_hw=(im_h, im_w), pool=POOLING_SIZE)

    # ---- heads ---------------------------------------------------------------------------------
    def _region_classification(self, spatial_fc7):
        """:277-290 -- 7x7 mean, the two Linears as one stacked GEMM, softmax / argmax (L2F.region_classification)."""
        cls_score, cls_pred, cls_prob, bbox_pred = L2F.region_classification(
            spatial_fc7, self.cls_score_net.weight, self.cls_score_net.bias, self.bbox_pred_net.weight,
            self.bbox_pred_net.bias)
        self._predictions["cls_score"] = cls_score
        self._predictions["cls_pred"] = cls_pred
        self._predictions["cls_prob"] = cls_prob
        self._predictions["bbox_pred"] = bbox_pred
        return cls_prob, bbox_pred

    def _mask_prediction(self, spatial_fc7, labels=None, mask_targets=None):
        """:292-307.  In training the reference already holds the mask targets when it predicts (they come from the
        proposal target layer, :586-590): passing them here makes prediction + mask loss (:404-413) one autograd node
        whose backward never materialises the (n,81,14,14) score gradient; `_mask_loss` then returns that loss."""
        ws = (self.mask_up_sampling.weight, self.mask_up_sampling.bias, self.mask_pred_net.weight, self.mask_pred_net.bias)
        if labels is not None and mask_targets is not None:
            mask_score, mask_prob, loss = L2F.mask_head_with_loss(spatial_fc7, *ws, labels, mask_targets)
            self._losses["mask_loss"] = loss
        else:
            mask_score, mask_prob = L2F.mask_head(spatial_fc7, *ws)
            self._losses.pop("mask_loss", None)
        self._predictions["mask_score"] = mask_score
        self._predictions["mask_prob"] = mask_prob
        return mask_prob

    # ---- dynamic filter ------------------------------------------------------------------------
    def _dynamic_filter(self, net_conv, labels=None, hidden=None, expr2img=None, resp_target=None, lengths=None,
                        cut_filters=False):
        """net_conv (I,C,H,W) ; labels (E,L) tokens or precomputed hidden (E,Dh).

        Stores _predictions['net_conv_before'] / ['response'] like :504,:568 and returns the gated map.
        cut_filters: the autograd graph is cut at the expression embedding (the filter generator sees a detached leaf;
        `_predictions["graph_cut"]` = ((hidden,), (leaf,))); the caller finishes the backward with
        `torch.autograd.backward([hidden], [leaf.grad])` -- bench.py (N > 1) runs that last part (the language encoder)
        while the gradient groups that are already complete are being all-reduced."""
        self._predictions["net_conv_before"] = net_conv
        if hidden is None:
            _, hidden, _ = self.rnn_encoder(labels, lengths)
        if cut_filters and torch.is_grad_enabled() and hidden.requires_grad:
            leaf = hidden.detach().requires_grad_(True)
            self._predictions["graph_cut"] = ((hidden,), (leaf,))
            hidden = leaf
        filt, fuse = generate_filters(hidden, [getattr(self, "dynamic_fc_%d" % k) for k in range(7)], self.response_fc)
        response, gated, resp_loss = L2F.dynamic_filter(net_conv, filt, fuse, expr2img, self._gate, resp_target)
        self._predictions["response"] = response
        self._predictions["net_conv"] = gated
        self._losses["loss_response_per_expr"] = resp_loss
        self._expr2img = expr2img.long() if torch.is_tensor(expr2img) else None
        return gated

    # ---- losses on the path ----------------------------------------------------------------------
    def _mask_loss(self, labels, mask_targets):
        if "mask_loss" in self._losses:          # computed together with the prediction
            return self._losses["mask_loss"]
        return L2F.mask_bce_loss(self._predictions["mask_score"], labels, mask_targets)

    def _caption_loss(self, fc_feats, att_feats, cap_labels, cap_masks, steps=None):
        """`steps`: host-known number of decode steps (max caption length + 1); None reads it from the labels."""
        return self.caption_model.forward_loss(fc_feats, att_feats, cap_labels, cap_masks, steps=steps)

    def _smooth_l1_loss(self, bbox_pred, bbox_targets, bbox_inside_weights, bbox_outside_weights, sigma=1.0, dim=(1,)):
        """:360-373, unchanged arithmetic (elementwise on (N, 4*classes): framework ops)."""
        sigma_2 = sigma ** 2
        in_box_diff = bbox_inside_weights * (bbox_pred - bbox_targets)
        abs_in_box_diff = torch.abs(in_box_diff)
        sign = (abs_in_box_diff < 1. / sigma_2).detach().float()
        in_loss_box = torch.pow(in_box_diff, 2) * (sigma_2 / 2.) * sign + (abs_in_box_diff - (0.5 / sigma_2)) * (1. - sign)
        loss_box = bbox_outside_weights * in_loss_box
        for i in sorted(dim, reverse=True):
            loss_box = loss_box.sum(i)
        return loss_box.mean()

    def _add_hot_path_losses(self, cap_labels=None, cap_masks=None, fc_feats=None, att_feats=None, steps=None,
                             rpn_cross_entropy=0., rpn_loss_box=0., num_expressions=1):
        """The hot-path terms of Network._add_losses (:375-453) from what the forward methods left in
        `_predictions` / `_proposal_targets` / `_losses`, under the reference's keys, and their weighted sum (:449):

            cross_entropy + loss_box + rpn_cross_entropy + rpn_loss_box + loss_mask + loss_response
                + cap_loss_weight * loss_caption

        The RPN terms belong to the caller (RPN is outside this repository's scope) and are passed in.  A term whose
        inputs are absent is skipped.  Batched extension: the reference runs one expression per step, so with
        `num_expressions` = E expressions in the batch (same number of ROIs each) every ROI-mean term is multiplied by
        E and the per-expression response losses are summed -- the total then equals the sum of the reference's E
        per-step losses (E = 1 reproduces :449 exactly).
        fc_feats / att_feats default to the caption features of `net_conv_before` / `net_conv` through
        `_head_to_tail` (:424-438) when a head was supplied."""
        E = float(num_expressions)
        L, P, T = self._losses, self._predictions, self._proposal_targets
        total = 0.
        if "cls_score" in P and "labels" in T:
            label = T["labels"].view(-1)
            score = P["cls_score"].reshape(-1, P["cls_score"].shape[-1])
            nll, _ = L2F.logsoftmax_nll(score, label, torch.ones_like(label, dtype=torch.float32))
            L["cross_entropy"] = E * nll / label.numel()
            total = total + L["cross_entropy"]
        if "bbox_pred" in P and "bbox_targets" in T:
            L["loss_box"] = E * self._smooth_l1_loss(P["bbox_pred"], T["bbox_targets"], T["bbox_inside_weights"],
                                                     T["bbox_outside_weights"])
            total = total + L["loss_box"]
        L["rpn_cross_entropy"], L["rpn_loss_box"] = rpn_cross_entropy, rpn_loss_box
        total = total + rpn_cross_entropy + rpn_loss_box
        if "mask_targets" in T and ("mask_loss" in L or "mask_score" in P):
            n_fg = T["mask_targets"].shape[0]
            L["loss_mask"] = E * self._mask_loss(T["labels"].view(-1)[:n_fg], T["mask_targets"])
            total = total + L["loss_mask"]
        if "loss_response_per_expr" in L:
            L["loss_response"] = L["loss_response_per_expr"].sum()
            total = total + L["loss_response"]
        if cap_labels is not None:
            if att_feats is None:
                after = self._head_to_tail(P["net_conv"])
                before = self._head_to_tail(P["net_conv_before"])       # once per IMAGE, then shared by its expressions
                if before.shape[0] != after.shape[0]:
                    before = before[self._expr2img]
                fc_feats, att_feats = L2F.caption_features(before, after, 14)
            # LanguageModelCriterion normalises by the number of caption tokens of the batch: for E > 1 the sum of the
            # per-expression losses is not a multiple of the batched mean, so the batched mean is used as is
            L["loss_caption"] = self._caption_loss(fc_feats, att_feats, cap_labels, cap_masks, steps=steps)
            total = total + self._cap_loss_weight * L["loss_caption"]
        L["total_loss"] = total
        return total


DEFAULT_OPT = dict(
    vocab_size=1999, word_embedding_size=512, word_vec_size=512, rnn_hidden_size=512, bidirectional=1,
    word_drop_out=0.5, rnn_drop_out=0.2, rnn_num_layers=1, rnn_type="lstm", variable_lengths=1,
    C4_feat_dim=1024, cap_loss_weight=1.0, caption_model="att2in2", input_encoding_size=512, rnn_size=512,
    num_layers=1, drop_prob_lm=0.5, seq_length=10, fc_feat_size=4096, att_feat_size=4096, att_hid_size=512,
)


class HotPathNet(Network):
    """Modules of resnetv1 (resnet_v1_cycle_response.py:232-335) that lie on the hot path."""

    def __init__(self, opt=None, num_classes=81, fc7_dim=2048, mask_mid=256, head_to_tail=None, batch_size=1):
        super().__init__(batch_size=batch_size)
        o = dict(DEFAULT_OPT)
        o.update(opt or {})
        self.opt = o
        self._num_classes = num_classes
        self.rnn_encoder = RNNEncoder(vocab_size=o["vocab_size"], word_embedding_size=o["word_embedding_size"],
                                      word_vec_size=o["word_vec_size"], hidden_size=o["rnn_hidden_size"],
                                      bidirectional=o["bidirectional"] > 0, input_dropout_p=o["word_drop_out"],
                                      dropout_p=o["rnn_drop_out"], n_layers=o["rnn_num_layers"],
                                      rnn_type=o["rnn_type"], variable_lengths=o["variable_lengths"] > 0)
        hid = o["rnn_num_layers"] * (2 if o["bidirectional"] > 0 else 1) * o["rnn_hidden_size"]
        self._C4_feat_dim = o["C4_feat_dim"]
        self._cap_loss_weight = o["cap_loss_weight"]
        self.caption_model = caption_models.setup(o)
        for k in range(7):
            setattr(self, "dynamic_fc_%d" % k, nn.Linear(hid, self._C4_feat_dim))
        self.response_fc = nn.Linear(hid, 7)
        self.cls_score_net = nn.Linear(fc7_dim, num_classes)
        self.bbox_pred_net = nn.Linear(fc7_dim, num_classes * 4)
        self.mask_up_sampling = nn.ConvTranspose2d(fc7_dim, mask_mid, 2, 2)
        self.mask_pred_net = nn.Conv2d(mask_mid, num_classes, kernel_size=1, stride=1)
        self._head = head_to_tail
        self.init_weights()

    def init_weights(self):
        # network_cycle_response.py:333-355: N(0, 0.01) heads (bbox 0.001), zero bias
        for m, std in ((self.cls_score_net, 0.01), (self.bbox_pred_net, 0.001), (self.mask_up_sampling, 0.01),
                       (self.mask_pred_net, 0.01)):
            m.weight.data.normal_(0, std)
            m.bias.data.zero_()

    def _head_to_tail(self, pool5):
        if self._head is None:
            raise RuntimeError("res5 / fc6-7 is outside this repository's scope: pass head_to_tail=<module>")
        return self._head(pool5)

    def chained_train_step(self, X, labels, expr2img, rois, roi_labels, gt_boxes, gt_masks, cap_labels, cap_masks,
                           num_fg, lengths=None, steps=None):
        """The reference-faithful TRAIN chain of `_predict` (:576-600) + `_add_losses` (:375-453) from the dynamic
        filter on, WITH res5 (`head_to_tail`) in the middle -- the "second number" of SURVEY 8d:

            dynamic filter (+ response target resized on the device, :418) -> 7x7 crop of the gated map -> res5 ->
            box head on every ROI (cross entropy + smooth L1) and mask head on the foreground ROIs (:592-596) ;
            ungated / gated map -> res5 -> caption features (:424-438) -> att2in2 + LM criterion ; weighted sum (:449).

        Batched layout: rois (N,5) [expression index,x1,y1,x2,y2] with the `num_fg` foreground ROIs of ALL expressions
        first (the reference keeps fg first, proposal_target_layer.py:177); roi_labels (N,) float (0 = background);
        gt_boxes (E,5) / gt_masks (E,imH,imW) uint8: one referred object per expression, so a ROI's assigned ground
        truth is its expression.  All targets (response, box, mask) are built on the device from these."""
        from ..layer_utils.proposal_target_layer import CFG as PT_CFG, bbox_regression_targets
        E, (H, W) = labels.shape[0], X.shape[2:]
        resp_tgt = L2F.resize_masks_nearest(gt_masks, H, W)
        gated = self._dynamic_filter(X, labels, expr2img=expr2img, resp_target=resp_tgt, lengths=lengths)
        pool5 = self._crop_pool_layer(gated, rois, max_pool=False)
        spatial_fc7 = self._head_to_tail(pool5)
        self._region_classification(spatial_fc7)
        assign = rois[:, 0].long()
        bt, biw = bbox_regression_targets(rois, gt_boxes, assign, roi_labels, self._num_classes, PT_CFG)
        mt = L2F.mask_targets(gt_masks, rois[:num_fg], assign[:num_fg], MASK_SIZE)
        self._proposal_targets = {"labels": roi_labels.long(), "bbox_targets": bt, "bbox_inside_weights": biw,
                                  "bbox_outside_weights": (biw > 0).float(), "mask_targets": mt}
        self._mask_prediction(spatial_fc7[:num_fg], roi_labels[:num_fg].long(), mt)
        return self._add_hot_path_losses(cap_labels, cap_masks, steps=steps, num_expressions=E)

    def gradient_groups(self):
        """The parameter groups whose gradients are all-reduced (north_star)."""
        fg = list(self.response_fc.parameters())
        for k in range(7):
            fg += list(getattr(self, "dynamic_fc_%d" % k).parameters())
        heads = [p for m in (self.cls_score_net, self.bbox_pred_net, self.mask_up_sampling, self.mask_pred_net)
                 for p in m.parameters()]
        # the filter generator of north_star / SURVEY 8e as two buckets: its projections (7,353,375 parameters, complete
        # as soon as the dynamic filter's backward has run) and the language encoder (5,489,640, the last gradients of
        # the step) -- so that only the encoder's 22 MB are all-reduced after the backward has finished
        groups = {"filter_generator": fg, "encoder": list(self.rnn_encoder.parameters()),
                  "caption": list(self.caption_model.parameters()), "heads": heads}
        if isinstance(self._head, nn.Module):      # trainable res5 glue (14.9 M more parameters)
            groups["res5"] = [p for p in self._head.parameters() if p.requires_grad]
        return groups
