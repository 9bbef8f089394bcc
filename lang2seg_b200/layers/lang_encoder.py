"""Expression encoder -- drop-in for lib/layers/lang_encoder.py:RNNEncoder.

Parameter names (embedding, mlp.0, rnn.*) and the returned triple match the reference so that its checkpoints
load (`self.rnn` is still an nn.LSTM: it owns the weights).  On the GPU the reference configuration (1 layer,
bidirectional LSTM, variable lengths) does not go through cuDNN: the input projection of all tokens and both
directions is one GEMM and the recurrence runs in l2s_bilstm_{fwd,bwd} with MASKING instead of packing -- no
host-side argsort (lang_encoder.py:38-52), no device->host read of the lengths, CUDA-graph capturable.  Other
configurations fall back to torch's LSTM with pack_padded_sequence(enforce_sorted=False).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

from .. import functional as L2F


class RNNEncoder(nn.Module):
    def __init__(self, vocab_size, word_embedding_size, word_vec_size, hidden_size, bidirectional=False,
                 input_dropout_p=0, dropout_p=0, n_layers=1, rnn_type="lstm", variable_lengths=True):
        super().__init__()
        self.variable_lengths = variable_lengths
        self.embedding = nn.Embedding(vocab_size, word_embedding_size)
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

        `lengths`: optional HOST tensor / list of the expression lengths.  The reference computes them on the host
        from the labels (lang_encoder.py:38-40, a device->host sync per call); passing them in keeps the launch
        queue asynchronous when the labels already live on the GPU."""
        B, L = input_labels.shape
        vec = self.mlp(self.input_dropout(self.embedding(input_labels)))
        if not self.variable_lengths:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                output, hidden = self.rnn(vec)
            return output, hidden, vec
        if (vec.is_cuda and self.rnn_type == "lstm" and self.rnn.num_layers == 1 and self.num_dirs == 2
                and self.rnn.hidden_size % 4 == 0):
            # masked recurrence: lengths stay on the device (the caller guarantees max length == L, lang_encoder.py:45)
            lens = lengths if (torch.is_tensor(lengths) and lengths.is_cuda) else (input_labels != 0).sum(1)
            r = self.rnn
            w_ih = torch.cat([r.weight_ih_l0, r.weight_ih_l0_reverse], 0)
            b = torch.cat([r.bias_ih_l0 + r.bias_hh_l0, r.bias_ih_l0_reverse + r.bias_hh_l0_reverse], 0)
            x2 = vec.reshape(B * L, -1)
            xg = L2F.linear(x2, w_ih, b) if (B * L >= 256 and x2.shape[1] % 8 == 0) else F.linear(x2, w_ih, b)
            output, hidden = L2F.bilstm(xg.view(B, L, -1), r.weight_hh_l0, r.weight_hh_l0_reverse, lens)
            keep = (torch.arange(L, device=vec.device)[None, :] < lens[:, None]).unsqueeze(-1)
            return output, hidden, vec * keep
        if lengths is None:
            lens_cpu = (input_labels != 0).sum(1).cpu()
        else:
            lens_cpu = torch.as_tensor(lengths, dtype=torch.int64, device="cpu")
        lengths = lens_cpu.to(vec.device, non_blocking=True)
        assert int(lens_cpu.max()) == L, "labels must be trimmed to the longest expression (lang_encoder.py:45)"
        packed = pack_padded_sequence(vec, lens_cpu, batch_first=True, enforce_sorted=False)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):     # fp32 parity: no TF32 inside cuDNN
            output, hidden = self.rnn(packed)
        output, _ = pad_packed_sequence(output, batch_first=True, total_length=L)
        if self.rnn_type == "lstm":
            hidden = hidden[0]
        hidden = hidden.transpose(0, 1).contiguous().view(B, -1)
        keep = (torch.arange(L, device=vec.device)[None, :] < lengths[:, None]).unsqueeze(-1)
        return output, hidden, vec * keep
