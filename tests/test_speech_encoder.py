"""The PyTorch encoder of config C5 (asr_chinese_e2e_b200/speech_encoder.py) against the reference's own Encoder
class: same parameter names (a reference state_dict loads unchanged) and same outputs.  Runs only where the reference
tree is mounted (this container); the GPU tests then use SpeechEncoder knowing it IS the reference's architecture."""
import pytest
import torch

from ref_import import HAVE_REFERENCE, mount_reference


@pytest.mark.skipif(not HAVE_REFERENCE, reason="reference tree only in the authoring container")
def test_matches_the_reference_encoder(monkeypatch):
    mount_reference(monkeypatch)
    from Predictor.Models.transformer_official import Encoder
    from asr_chinese_e2e_b200.speech_encoder import SpeechEncoder
    torch.manual_seed(0)
    kw = dict(d_input=24, n_layers=3, n_head=4, d_k=8, d_v=8, d_model=32, d_inner=64, dropout=0.1)
    ref = Encoder(**kw).eval()
    mine = SpeechEncoder(**kw).eval()
    missing = mine.load_state_dict(ref.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    x = torch.randn(5, 17, 24)
    lens = torch.tensor([17, 9, 17, 3, 12])
    x = x * (torch.arange(17)[None, :, None] < lens[:, None, None])
    with torch.no_grad():
        a, = ref(x, lens)
        b, = mine(x, lens)
    assert a.shape == b.shape == (5, 17, 32)
    assert torch.allclose(a, b, atol=2e-5), (a - b).abs().max()
    assert torch.all(b[1, 9:] == 0) and torch.all(b[3, 3:] == 0)          # padded frames leave as zeros
    # gradients agree as well (train mode would differ only by the dropout RNG stream)
    for m in (ref, mine):
        m.zero_grad()
        m(x, lens)[0].pow(2).sum().backward()
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), mine.named_parameters()):
        assert n1 == n2 and torch.allclose(p1.grad, p2.grad, atol=1e-3, rtol=1e-3), n1
