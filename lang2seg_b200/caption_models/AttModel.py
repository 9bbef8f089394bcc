"""att2in2 caption model on the sm_100a kernels -- drop-in for lib/caption_models/AttModel.py.

Classes, constructor options, parameter names and call signatures follow the reference
(AttModel :27-110, Attention :397-423, Att2in2Core :426-466, Att2in2Model :479-484) so that
`caption_model.core.attention.h2att.weight` etc. load from its checkpoints.  What changes is where
the arithmetic runs:

  Attention.forward    h2att Linear (skinny exact-fp32 GEMM) + ONE fused kernel: score, softmax, weighted sum
  Att2in2Core.forward  i2h/h2h/a2c Linears (skinny exact-fp32 GEMM) + one gate kernel (sigmoid/maxout/tanh/cell)
  AttModel.forward     same T-step loop with the reference's early break; `forward_loss` fuses
                       logit -> log-softmax -> masked NLL (LanguageModelCriterion) per step
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import functional as L2F


class Attention(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.rnn_size = opt["rnn_size"]
        self.att_hid_size = opt["att_hid_size"]
        self.h2att = nn.Linear(self.rnn_size, self.att_hid_size)
        self.alpha_net = nn.Linear(self.att_hid_size, 1)

    def forward(self, h, att_feats, p_att_feats):
        B = att_feats.size(0)
        att_size = att_feats.numel() // B // self.rnn_size
        att_h = L2F.dense(h, self.h2att.weight, self.h2att.bias)
        res, _ = L2F.attention_step(att_h, att_feats.reshape(B, att_size, self.rnn_size),
                                    p_att_feats.reshape(B, att_size, self.att_hid_size),
                                    self.alpha_net.weight, self.alpha_net.bias)
        return res


class Att2in2Core(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.input_encoding_size = opt["input_encoding_size"]
        self.rnn_size = opt["rnn_size"]
        self.drop_prob_lm = opt["drop_prob_lm"]
        self.fc_feat_size = opt["fc_feat_size"]
        self.att_feat_size = opt["att_feat_size"]
        self.att_hid_size = opt["att_hid_size"]
        self.a2c = nn.Linear(self.rnn_size, 2 * self.rnn_size)
        self.i2h = nn.Linear(self.input_encoding_size, 5 * self.rnn_size)
        self.h2h = nn.Linear(self.rnn_size, 5 * self.rnn_size)
        self.dropout = nn.Dropout(self.drop_prob_lm)
        self.attention = Attention(opt)

    def forward(self, xt, fc_feats, att_feats, p_att_feats, state):
        h_prev, c_prev = state[0][-1], state[1][-1]
        att_res = self.attention(h_prev, att_feats, p_att_feats)
        sums = L2F.dense(xt, self.i2h.weight, self.i2h.bias) + L2F.dense(h_prev, self.h2h.weight, self.h2h.bias)
        next_h, next_c = L2F.att2in2_gates(sums, L2F.dense(att_res, self.a2c.weight, self.a2c.bias), c_prev)
        output = self.dropout(next_h)
        return output, (next_h.unsqueeze(0), next_c.unsqueeze(0))


class AttModel(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.vocab_size = opt["vocab_size"]
        self.input_encoding_size = opt["input_encoding_size"]
        self.rnn_size = opt["rnn_size"]
        self.num_layers = opt["num_layers"]
        self.drop_prob_lm = opt["drop_prob_lm"]
        self.seq_length = opt["seq_length"]
        self.fc_feat_size = opt["fc_feat_size"]
        self.att_feat_size = opt["att_feat_size"]
        self.att_hid_size = opt["att_hid_size"]
        self.ss_prob = 0.0
        self.embed = nn.Sequential(L2F.Embedding(self.vocab_size + 1, self.input_encoding_size), nn.ReLU(),
                                   nn.Dropout(self.drop_prob_lm))
        self.fc_embed = nn.Sequential(nn.Linear(self.fc_feat_size, self.rnn_size), nn.ReLU(),
                                      nn.Dropout(self.drop_prob_lm))
        self.att_embed = nn.Sequential(nn.Linear(self.att_feat_size, self.rnn_size), nn.ReLU(),
                                       nn.Dropout(self.drop_prob_lm))
        self.logit = nn.Linear(self.rnn_size, self.vocab_size + 1)
        self.ctx2att = nn.Linear(self.rnn_size, self.att_hid_size)

    def init_hidden(self, bsz):
        w = next(self.parameters())
        return (w.new_zeros(self.num_layers, bsz, self.rnn_size), w.new_zeros(self.num_layers, bsz, self.rnn_size))

    @staticmethod
    def decode_steps(seq):
        """T of the reference loop: stop at the first column i >= 1 that is all zero (AttModel.py:92-93)."""
        cols = (seq[:, 1:-1] != 0).any(0) if seq.size(1) > 2 else seq.new_zeros(0, dtype=torch.bool)
        nz = torch.nonzero(~cols)
        k = int(nz[0]) if nz.numel() else cols.numel()
        return min(seq.size(1) - 1, 1 + k)

    @staticmethod
    def _big_linear(lin, x):
        """The (B*196)-row projections dominate the caption model's flops (822 MFLOP/sample): L2F.dense runs them on
        the tcgen05 bf16x3 GEMM; small batches take the skinny exact-fp32 GEMM.  No cuBLAS behind either."""
        return L2F.dense(x, lin.weight, lin.bias)

    def _prepare(self, fc_feats, att_feats):
        fc_feats = self.fc_embed(fc_feats)
        att = self._big_linear(self.att_embed[0], att_feats.reshape(-1, self.att_feat_size))
        att = self.att_embed[2](self.att_embed[1](att))                      # ReLU, Dropout (AttModel.py:49-51)
        att = att.view(*(att_feats.size()[:-1] + (self.rnn_size,)))
        p_att = self._big_linear(self.ctx2att, att.view(-1, self.rnn_size))
        p_att = p_att.view(*(att.size()[:-1] + (self.att_hid_size,)))
        return fc_feats, att, p_att

    def _fast_decode_ok(self, att):
        core = self.core
        if not att.is_cuda:
            raise L2F._lib.L2SError("AttModel: CUDA tensors only (there is no CPU path; the CPU baseline is oracle/)")
        return (isinstance(core, Att2in2Core) and self.num_layers == 1
                and self.rnn_size == self.att_hid_size and self.rnn_size % 4 == 0 and self.rnn_size <= 1024)

    def _decode(self, att, p_att, seq, T):
        """All T steps of the recurrence through l2s_att2in2_decode_{fwd,bwd}: i2h(x_t) for every t is one GEMM,
        the state-dependent part is four launches per token inside the library.  Returns dropout(h_t) (T,B,D)."""
        core = self.core
        B = att.size(0)
        xt = self.embed(seq[:, :T].t())                                   # (T,B,E): Embedding+ReLU+Dropout (:95)
        bias = core.i2h.bias + core.h2h.bias
        i2h_all = L2F.dense(xt.reshape(T * B, -1), core.i2h.weight, bias)
        att3 = att.reshape(B, -1, self.rnn_size)
        h_all = L2F.att2in2_decode(i2h_all.view(T, B, -1), att3, p_att.reshape(B, -1, self.att_hid_size),
                                   core.attention.h2att.weight, core.attention.h2att.bias, core.h2h.weight,
                                   core.a2c.weight, core.a2c.bias, core.attention.alpha_net.weight,
                                   core.attention.alpha_net.bias)
        return core.dropout(h_all)

    def forward(self, fc_feats, att_feats, seq, steps=None):
        """fc_feats (B,F) ; att_feats (B,14,14,F) ; seq (B,L+2) -> log-probs (B,T,V+1).
        `steps`: host-known T (skips the device->host read of decode_steps)."""
        B = fc_feats.size(0)
        fc_feats, att, p_att = self._prepare(fc_feats, att_feats)
        T = int(steps) if steps is not None else self.decode_steps(seq)
        if self._fast_decode_ok(att):
            out = self._decode(att, p_att, seq, T)                        # (T,B,D)
            logits = self._big_linear(self.logit, out.reshape(T * B, -1))
            logp = F.log_softmax(logits, dim=1) if torch.is_grad_enabled() else L2F.log_softmax(logits)
            return logp.view(T, B, -1).transpose(0, 1)
        state = self.init_hidden(B)
        outputs = []
        for i in range(T):
            xt = self.embed(seq[:, i])
            output, state = self.core(xt, fc_feats, att, p_att, state)
            logits = self._big_linear(self.logit, output)
            outputs.append(L2F.log_softmax(logits) if not torch.is_grad_enabled() else F.log_softmax(logits, dim=1))
        return torch.stack(outputs, 1)

    def forward_loss(self, fc_feats, att_feats, seq, masks, steps=None):
        """crit(model(fc, att, seq), seq[:,1:], masks[:,1:]) of network_cycle_response.py:443 without
        materialising the (B,T,V+1) log-probs: logit for all steps -> fused log-softmax + masked NLL."""
        B = fc_feats.size(0)
        fc_feats, att, p_att = self._prepare(fc_feats, att_feats)
        T = int(steps) if steps is not None else self.decode_steps(seq)
        masks = masks.to(att.dtype)
        if self._fast_decode_ok(att):
            out = self._decode(att, p_att, seq, T)
            logits = self._big_linear(self.logit, out.reshape(T * B, -1))
            nll, _ = L2F.logsoftmax_nll(logits, seq[:, 1:T + 1].t().reshape(-1), masks[:, 1:T + 1].t().reshape(-1))
            return nll / masks[:, 1:T + 1].sum()
        state = self.init_hidden(B)
        total = att.new_zeros(())
        for i in range(T):
            xt = self.embed(seq[:, i])
            output, state = self.core(xt, fc_feats, att, p_att, state)
            nll, _ = L2F.logsoftmax_nll(self._big_linear(self.logit, output), seq[:, i + 1], masks[:, i + 1])
            total = total + nll
        return total / masks[:, 1:T + 1].sum()

    def get_logprobs_state(self, it, tmp_fc_feats, tmp_att_feats, tmp_p_att_feats, state):
        xt = self.embed(it)
        output, state = self.core(xt, tmp_fc_feats, tmp_att_feats, tmp_p_att_feats, state)
        return L2F.log_softmax(self._big_linear(self.logit, output)), state


class Att2in2Model(AttModel):
    def __init__(self, opt):
        super().__init__(opt)
        self.core = Att2in2Core(opt)
        delattr(self, "fc_embed")
        self.fc_embed = lambda x: x
