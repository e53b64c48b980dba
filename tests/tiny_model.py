"""A tiny encoder/decoder with the reference model's call signatures (encoder(wave, wave_len) -> (enc, ...),
decoder(tgt, enc, lens) -> (pred, gold, ...)) for testing the JointCTCAttention mix-in on CPU and GPU."""
import torch

from asr_chinese_e2e_b200 import JointCTCAttention, Pack


class TinyEnc(torch.nn.Module):
    def __init__(self, d_in, d):
        super().__init__()
        self.p = torch.nn.Linear(d_in, d)

    def forward(self, wave, wave_len):
        mask = (torch.arange(wave.size(1), device=wave.device)[None, :] < wave_len[:, None]).unsqueeze(-1).float()
        return (torch.tanh(self.p(wave)) * mask,)      # padded frames zeroed like the reference encoder


class TinyDec(torch.nn.Module):
    def __init__(self, d, V):
        super().__init__()
        self.emb, self.out = torch.nn.Embedding(V, d), torch.nn.Linear(d, V)

    def forward(self, tgt, enc, lens):
        B = tgt.size(0)
        ys = torch.cat([torch.full((B, 1), 2, device=tgt.device), tgt], 1)                  # <sos> + tokens
        gold = torch.cat([tgt, torch.zeros(B, 1, dtype=torch.long, device=tgt.device)], 1)
        gold[torch.arange(B, device=tgt.device), lens] = 3               # <eos>
        return self.out(self.emb(ys) + enc.mean(1, keepdim=True)), gold


class TinyJoint(JointCTCAttention, torch.nn.Module):
    def __init__(self, V=13, d=16):
        torch.nn.Module.__init__(self)
        self.encoder, self.decoder = TinyEnc(8, d), TinyDec(d, V)
        self.init_ctc(d, V, ctc_weight=0.3, ctc_zero_infinity=True)


def _batch(B=3, T=12, U=4, V=13):
    g = torch.Generator().manual_seed(5)
    tl = torch.tensor([4, 2, 3])
    tg = torch.randint(4, V, (B, U), generator=g) * (torch.arange(U)[None] < tl[:, None])
    return Pack(wave=torch.randn(B, T, 8, generator=g), wave_len=torch.tensor([12, 9, 7]),
                tgt_for_input=tg, tgt_len=tl)
