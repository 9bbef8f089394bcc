"""Expression encoder -- drop-in for lib/layers/lang_encoder.py:RNNEncoder.

Stays on torch (cuDNN LSTM): it is the input of the filter generator, not one of the hot kernels.
Parameter names (embedding, mlp.0, rnn) and the returned triple match the reference so that its
checkpoints load.  Differences: sorting by length is left to pack_padded_sequence
(enforce_sorted=False) instead of the host-side numpy argsort at lang_encoder.py:38-52.
"""
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence


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
