"""Developer tool (GPU box): BASELINE.json config 5 -- a 12-layer speech-Transformer encoder forward/backward in
plain PyTorch feeding the CTC kernels, B=128, T=400 -- to put the CTC path in the context of a training step.

The encoder re-states the reference's architecture for timing only (Predictor/Models/transformer_official.py:
128-213: Linear(320->512)+LayerNorm+sinusoidal PE, N x [8-head self-attention with post-LN residual, Conv1d(k=1)
FFN 512->1024->512 with post-LN residual, padded frames zeroed]); it is not part of the product.
Compares the step with (a) this repo's op and (b) torch's own CUDA log_softmax + ctc_loss."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
from asr_chinese_e2e_b200 import ctc_loss_b200
from oracle.synth import make_targets, make_lengths

B, T, V, U, D, H, FF, NL = 128, 400, 4234, 50, 512, 8, 1024, int(os.environ.get("LAYERS", "12"))


class Layer(nn.Module):
    def __init__(self):
        super().__init__()
        self.qkv = nn.Linear(D, 3 * D); self.fc = nn.Linear(D, D); self.ln1 = nn.LayerNorm(D)
        self.w1 = nn.Conv1d(D, FF, 1); self.w2 = nn.Conv1d(FF, D, 1); self.ln2 = nn.LayerNorm(D)

    def forward(self, x, keep, attn_mask):
        b, t, _ = x.shape
        q, k, v = self.qkv(x).view(b, t, 3, H, D // H).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=attn_mask)
        x = self.ln1(self.fc(a.transpose(1, 2).reshape(b, t, D)) + x) * keep
        y = self.w2(F.relu(self.w1(x.transpose(1, 2)))).transpose(1, 2)
        return self.ln2(y + x) * keep


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.inp = nn.Linear(320, D); self.ln = nn.LayerNorm(D)
        self.layers = nn.ModuleList([Layer() for _ in range(NL)])
        pe = torch.zeros(T, D); pos = torch.arange(T).unsqueeze(1).float()
        div = torch.exp(torch.arange(0, D, 2).float() * -(math.log(10000.0) / D))
        pe[:, 0::2] = torch.sin(pos * div); pe[:, 1::2] = torch.cos(pos * div)
        self.register_buffer("pe", pe)
        self.head = nn.Linear(D, V)

    def forward(self, wave, wave_len):
        t = wave.size(1)
        valid = torch.arange(t, device=wave.device)[None, :] < wave_len[:, None]
        keep = valid.unsqueeze(-1).float()
        mask = valid[:, None, None, :]
        x = self.ln(self.inp(wave)) + self.pe[:t]
        for l in self.layers:
            x = l(x, keep, mask)
        return self.head(x)


def main():
    g = torch.Generator().manual_seed(1005)
    tg, tl = make_targets(B, U, V, g); il = make_lengths(B, T, g)
    wave = torch.randn(B, T, 320, generator=g).cuda(); tg, tl, il = tg.cuda(), tl.cuda(), il.cuda()
    m = Encoder().cuda()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)

    def step(kind):
        opt.zero_grad(set_to_none=True)
        logits = m(wave, il)
        if kind == "b200":
            loss = ctc_loss_b200(logits, tg, il, tl, zero_infinity=True)
        elif kind == "torch":
            loss = F.ctc_loss(F.log_softmax(logits, -1).transpose(0, 1), tg, il, tl, zero_infinity=True)
        else:
            loss = logits.float().pow(2).mean()          # encoder + head only (no CTC): lower bound of the step
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 5.0)
        opt.step()
        return loss

    res = {}
    for kind in ("none", "b200", "torch"):
        for _ in range(3): step(kind)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): loss = step(kind)
        e1.record(); torch.cuda.synchronize()
        res[kind] = e0.elapsed_time(e1) / 10
        print(f"{kind:6s}: {res[kind]:8.2f} ms/step  loss {loss.item():.4f}")
    print(f"CTC share of the joint step: this repo {(res['b200'] - res['none']) / res['b200'] * 100:.1f} %  "
          f"({res['b200'] - res['none']:.2f} ms), torch CUDA ctc {(res['torch'] - res['none']) / res['torch'] * 100:.1f} % "
          f"({res['torch'] - res['none']:.2f} ms); layers={NL} B={B} T={T} fp32 (TF32 off)")


if __name__ == "__main__":
    main()
