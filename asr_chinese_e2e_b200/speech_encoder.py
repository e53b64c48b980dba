"""The caller on the input side of the CTC path: the reference's speech-Transformer ENCODER, re-stated in PyTorch
for BASELINE.json config C5 ("12-layer Transformer encoder forward in PyTorch feeding the CTC kernels").

Architecture and parameter names follow ``Encoder`` / ``EncoderLayer`` of the reference
(Predictor/Models/transformer_official.py:126-213) with ``MultiHeadAttention`` (Predictor/Models/attention.py:6-62)
and ``PositionwiseFeedForwardUseConv`` (Predictor/Models/module.py:58-75): Linear(d_input -> d_model) + LayerNorm +
sinusoidal PE, then n_layers x [n_head-way self-attention with post-LN residual, Conv1d(k=1) feed-forward with
post-LN residual], every sub-layer output multiplied by the non-pad mask (so padded frames leave the encoder as
zeros and the CTC head sees its bias there: SURVEY.md 8a-a3).  A ``state_dict`` of the reference's encoder loads
into this module unchanged (tests/test_speech_encoder.py checks outputs against the real class when the reference
tree is present).  What differs is the plumbing, not the arithmetic:

  * masks come from ``masks.py`` (one broadcast compare on the device; the reference loops over the batch on the
    host, Predictor/Models/utils.py:100-127);
  * attention runs through ``F.scaled_dot_product_attention`` with a key-padding mask instead of materialising
    ``(n_head*B) x T x T`` scores, soft-max and dropout tensors.

This is PyTorch plumbing (cuBLAS / SDPA): not one of this repo's kernels, and not on the CTC hot path itself.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .masks import get_non_pad_mask


class _PE(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        pos = torch.arange(0, max_len).unsqueeze(1).float()
        div = torch.exp(torch.arange(0, d_model, 2).float() * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(pos * div)
        pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        return self.pe[:, : x.size(1)]


class _SelfAttention(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout):
        super().__init__()
        self.n_head, self.d_k, self.d_v, self.p = n_head, d_k, d_v, dropout
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        nn.init.normal_(self.w_qs.weight, mean=0, std=math.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_ks.weight, mean=0, std=math.sqrt(2.0 / (d_model + d_k)))
        nn.init.normal_(self.w_vs.weight, mean=0, std=math.sqrt(2.0 / (d_model + d_v)))
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        nn.init.xavier_normal_(self.fc.weight)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, key_ok):
        b, t, _ = x.shape
        h = self.n_head
        q = self.w_qs(x).view(b, t, h, self.d_k).transpose(1, 2)
        k = self.w_ks(x).view(b, t, h, self.d_k).transpose(1, 2)
        v = self.w_vs(x).view(b, t, h, self.d_v).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=key_ok, dropout_p=self.p if self.training else 0.0,
                                           scale=1.0 / math.sqrt(self.d_k))
        o = o.transpose(1, 2).reshape(b, t, h * self.d_v)
        return self.layer_norm(self.dropout(self.fc(o)) + x)


class _ConvFFN(nn.Module):
    def __init__(self, d_in, d_hid, dropout):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, 1)
        self.w_2 = nn.Conv1d(d_hid, d_in, 1)
        self.layer_norm = nn.LayerNorm(d_in)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        y = self.w_2(F.relu(self.w_1(x.transpose(1, 2)))).transpose(1, 2)
        return self.layer_norm(self.dropout(y) + x)


class _EncoderLayer(nn.Module):
    def __init__(self, d_model, d_inner, n_head, d_k, d_v, dropout):
        super().__init__()
        self.slf_attn = _SelfAttention(n_head, d_model, d_k, d_v, dropout)
        self.pos_ffn = _ConvFFN(d_model, d_inner, dropout)

    def forward(self, x, keep, key_ok):
        x = self.slf_attn(x, key_ok) * keep
        return self.pos_ffn(x) * keep


class SpeechEncoder(nn.Module):
    """``Encoder(d_input, n_layers, n_head, d_k, d_v, d_model, d_inner, dropout)`` of the reference; defaults are the
    reference's (transformer_official.py:41-52,112-124: d_input = n_mels*lfr_m = 320, d_model 512, 8 heads x 64,
    FFN 1024, dropout 0.1), except n_layers which callers set (6 in the reference, 12 in config C5)."""

    def __init__(self, d_input=320, n_layers=6, n_head=8, d_k=64, d_v=64, d_model=512, d_inner=1024, dropout=0.1,
                 pe_maxlen=5000):
        super().__init__()
        self.d_model = d_model
        self.linear_in = nn.Linear(d_input, d_model)
        self.layer_norm_in = nn.LayerNorm(d_model)
        self.positional_encoding = _PE(d_model, pe_maxlen)
        self.dropout = nn.Dropout(dropout)
        self.layer_stack = nn.ModuleList([_EncoderLayer(d_model, d_inner, n_head, d_k, d_v, dropout)
                                          for _ in range(n_layers)])
        nn.init.xavier_normal_(self.linear_in.weight)

    def forward(self, padded_input, input_lengths, return_attns=False):
        keep = get_non_pad_mask(padded_input, input_lengths=input_lengths)          # [B,T,1] float
        key_ok = keep.squeeze(-1).bool()[:, None, None, :]                           # attend to valid frames only
        x = self.dropout(self.layer_norm_in(self.linear_in(padded_input)) + self.positional_encoding(padded_input))
        for layer in self.layer_stack:
            x = layer(x, keep, key_ok)
        return (x,)
