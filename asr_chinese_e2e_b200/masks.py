"""Length / padding masks without host loops.

Same names, arguments and results as the reference's helpers in ``Predictor/Models/utils.py:99-145`` (also
``common_layers.py:29-67``), which ``TransformerOffical``'s encoder and decoder call on every step
(``transformer_official.py:170-172, 292-303``).  The reference builds the length mask with
``for i in range(N): mask[i, input_lengths[i]:] = 0`` -- N slice assignments, and with lengths on the GPU
N device->host syncs -- per encoder call and again per decoder call.  Here it is one broadcast compare on the
device the data lives on; nothing is read back to the host.
"""
import torch

__all__ = ["get_non_pad_mask", "get_subsequent_mask", "get_attn_key_pad_mask", "get_attn_pad_mask"]


def _lengths_on(device, input_lengths):
    if not torch.is_tensor(input_lengths):
        input_lengths = torch.as_tensor(list(input_lengths), dtype=torch.int64)
    return input_lengths.to(device=device, non_blocking=True).reshape(-1)


def get_non_pad_mask(padded_input, input_lengths=None, pad_idx=None):
    """Padding positions 0, others 1; ``[N, T, 1]`` (reference utils.py:99-115).

    With ``input_lengths``: ``padded_input`` is ``[N, T, ...]`` and the mask has its dtype (``new_ones``).
    With ``pad_idx``: ``padded_input`` is ``[N, T]`` token ids and the mask is float32.  If both are given the
    ``pad_idx`` form wins, as in the reference."""
    assert input_lengths is not None or pad_idx is not None
    if pad_idx is not None:
        assert padded_input.dim() == 2
        return padded_input.ne(pad_idx).float().unsqueeze(-1)
    n, t = padded_input.shape[0], padded_input.shape[1]
    lens = _lengths_on(padded_input.device, input_lengths)
    assert lens.numel() == n
    # a negative length slices from the end in the reference (mask[i, -k:] = 0): same rule here
    lens = torch.where(lens < 0, (lens + t).clamp(min=0), lens)
    pos = torch.arange(t, device=padded_input.device)
    return (pos.unsqueeze(0) < lens.unsqueeze(1)).to(padded_input.dtype).unsqueeze(-1)


def get_subsequent_mask(seq):
    """``[B, L, L]`` uint8, 1 strictly above the diagonal (reference utils.py:117-125)."""
    sz_b, len_s = seq.size()
    pos = torch.arange(len_s, device=seq.device)
    mask = (pos.unsqueeze(0) > pos.unsqueeze(1)).to(torch.uint8)
    return mask.unsqueeze(0).expand(sz_b, -1, -1)


def get_attn_key_pad_mask(seq_k, seq_q, pad_idx):
    """``[B, Lq, Lk]`` bool, True where the key is padding (reference utils.py:127-135)."""
    return seq_k.eq(pad_idx).unsqueeze(1).expand(-1, seq_q.size(1), -1)


def get_attn_pad_mask(padded_input, input_lengths, expand_length):
    """``[N, expand_length, T]`` bool, True at padded key positions (reference utils.py:137-145)."""
    non_pad = get_non_pad_mask(padded_input, input_lengths=input_lengths)
    return non_pad.squeeze(-1).lt(1).unsqueeze(1).expand(-1, expand_length, -1)
