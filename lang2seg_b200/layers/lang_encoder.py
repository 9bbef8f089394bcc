"""Expression encoder -- drop-in for lib/layers/lang_encoder.py:RNNEncoder.

Parameter names (embedding, mlp.0, rnn.*) and the returned triple match the reference so that its checkpoints
load (`self.rnn` is still an nn.LSTM: it owns the weights).  On the GPU the reference configuration (1 layer,
bidirectional LSTM, variable lengths) does not go through cuDNN: the input projection of all tokens and both
directions is one GEMM and the recurrence runs in l2s_bilstm_{fwd,bwd} with MASKING instead of packing -- no
host-side argsort (lang_encoder.py:38-52), no device->host read of the lengths, CUDA-graph capturable.  Other
configurations (GRU, several layers, unidirectional, fixed lengths) are not on lang2seg's path and raise: there is no
cuDNN or CPU path behind this module.
"""
import torch
import torch.nn as nn

from .. import functional as L2F


class RNNEncoder(nn.Module):
    def __init__(self, vocab_size, word_embedding_size, word_vec_size, hidden_size, bidirectional=False,
                 input_dropout_p=0, dropout_p=0, n_layers=1, rnn_type="lstm", variable_lengths=True):
        super().__init__()
        self.variable_lengths = variable_lengths
        self.embedding = L2F.Embedding(vocab_size, word_embedding_size)
        self.input_dropout = nn.Dropout(input_dropout_p)
        self.mlp = nn.Sequential(nn.Linear(word_embedding_size, word_vec_size), nn.ReLU())
        self.rnn_type = rnn_type
        self.rnn = getattr(nn, rnn_type.upper())(word_vec_size, hidden_size, n_layers, batch_first=True,
                                                 bidirectional=bidirectional,
                                                 dropout=dropout_p if n_layers > 1 else 0)
        self.num_dirs = 2 if bidirectional else 1

    def train(self, mode=True):
        # cuDNN refuses RNN backward in eval mode; with one layer the LSTM has no dropout, so keeping the
        # LSTM itself in training mode changes nothing numerically (parity runs use eval())
        super().train(mode)
        if self.rnn.num_layers == 1:
            self.rnn.train(True)
        return self

    def forward(self, input_labels, lengths=None):
        """input_labels (B,L) int64 zero padded -> output (B,L,H*dirs), hidden (B,layers*dirs*H), embedded (B,L,Dw)

        `lengths`: optional DEVICE tensor of the expression lengths (the reference computes them on the host from the
        labels, lang_encoder.py:38-40: a device->host sync per call; here they never leave the device).

        Only the configuration lang2seg trains with (tools/opt.py: one layer, bidirectional LSTM, variable lengths) is
        on the hot path and implemented; anything else raises -- there is no cuDNN / CPU path behind this module."""
        if not input_labels.is_cuda:
            raise L2F._lib.L2SError("RNNEncoder: CUDA tensors only (there is no CPU path; the CPU baseline is oracle/)")
        if not (self.variable_lengths and self.rnn_type == "lstm" and self.rnn.num_layers == 1 and self.num_dirs == 2
                and self.rnn.hidden_size % 4 == 0):
            raise NotImplementedError("RNNEncoder: only the reference's configuration (1-layer bidirectional LSTM, "
                                      "variable lengths, hidden size % 4 == 0) is implemented on the B200 path")
        B, L = input_labels.shape
        emb = self.input_dropout(self.embedding(input_labels))
        vec = torch.relu(L2F.dense(emb, self.mlp[0].weight, self.mlp[0].bias))        # mlp = Linear + ReLU (:31)
        # masked recurrence: lengths stay on the device (the caller guarantees max length == L, lang_encoder.py:45)
        lens = lengths if (torch.is_tensor(lengths) and lengths.is_cuda) else (input_labels != 0).sum(1)
        r = self.rnn
        w_ih = torch.cat([r.weight_ih_l0, r.weight_ih_l0_reverse], 0)
        b = torch.cat([r.bias_ih_l0 + r.bias_hh_l0, r.bias_ih_l0_reverse + r.bias_hh_l0_reverse], 0)
        xg = L2F.dense(vec.reshape(B * L, -1), w_ih, b)
        output, hidden = L2F.bilstm(xg.view(B, L, -1), r.weight_hh_l0, r.weight_hh_l0_reverse, lens)
        keep = (torch.arange(L, device=vec.device)[None, :] < lens[:, None]).unsqueeze(-1)
        return output, hidden, vec * keep
