"""Developer tool (GPU box): per-utterance comparison of the lattice kernel with the float64 oracle for one
fuzz case, both recursions (CTCB200_LATTICE_LOG=1 in a second process to force log space)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle.c_oracle import ctc_c_f64
from oracle.synth import make_case
sys.path.insert(0, "tests")
from test_gpu_parity import _raw_loss_grad

B, T, V, umax, seed, dist = [int(a) if a.isdigit() else a for a in sys.argv[1:7]]
c = make_case(B, T, V, umax, seed, dist=dist, n_infeasible=1 if B > 2 and umax > 0 else 0,
              n_partial=1 if B > 3 and umax > 1 else 0)
if B > 1 and umax > 0:
    c["target_lengths"][-1] = 0
    c["targets"][-1] = 0
c["input_lengths"].clamp_(max=T)
nll, g, stats = _raw_loss_grad(c)
_, n64, g64 = ctc_c_f64(*[c[k].numpy() for k in ("logits", "targets", "input_lengths", "target_lengths")],
                        reduction="sum", zero_infinity=True)
print("stats", stats)
for b in range(B):
    print(b, "Tb", int(c["input_lengths"][b]), "Ub", int(c["target_lengths"][b]), "nll", nll[b], "f64", n64[b],
          "err", nll[b] - n64[b], "gerr", np.abs(g[b] - g64[b]).max())
