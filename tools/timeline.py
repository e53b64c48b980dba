"""Developer tool (GPU box): event timeline of the chunked two-stream pipeline (full-length C2 batch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_chinese_e2e_b200 import _lib
from oracle.synth import make_case

B, T, V, U = 256, 400, 4234, 50
n_ch = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = make_case(B, T, V, U, 1002, full_lengths=True)
x = c["logits"].cuda(); tg = c["targets"].cuda(); il = c["input_lengths"].cuda(); tl = c["target_lengths"].cuda()
L = _lib.lib()
per = B // n_ch
wsb = _lib.workspace_bytes(per, T, V, U)
ws = torch.empty(n_ch * wsb, dtype=torch.uint8, device="cuda")
nll = torch.empty(B, device="cuda"); sums = torch.zeros(n_ch, 4, device="cuda"); grad = torch.empty_like(x)
side = (torch.cuda.Stream(), torch.cuda.Stream())
ev = lambda: torch.cuda.Event(enable_timing=True)
def run(record):
    main = torch.cuda.current_stream()
    e0 = ev(); e0.record(main)
    for s in side: s.wait_event(e0)
    marks = []
    prev = None
    for ci in range(n_ch):
        lo = ci * per
        s = side[ci & 1]
        if prev is not None: s.wait_event(prev)
        es0 = ev(); es0.record(s)
        sw = ev(); sw.record(s)   # materialise handle
        rc = L.ctcb200_loss_grad(x.data_ptr() + lo*T*V*4, tg.data_ptr() + lo*U*8, U, tg.numel() - lo*U, il.data_ptr()+lo*8, tl.data_ptr()+lo*8,
                                 per, T, V, U, 0, 0, 1, 1.0/B, nll.data_ptr()+lo*4, sums.data_ptr()+ci*16, grad.data_ptr()+lo*T*V*4,
                                 ws.data_ptr()+ci*wsb, wsb, s.cuda_stream, sw.cuda_event)
        assert rc == 0
        ee = ev(); ee.record(s)
        marks.append((es0, sw, ee)); prev = sw
    for s in side:
        e = ev(); e.record(s); main.wait_event(e)
    e1 = ev(); e1.record(main)
    torch.cuda.synchronize()
    if record:
        print(f"chunks={n_ch} total {e0.elapsed_time(e1)*1e3:.0f} us")
        for i, (a, b_, c_) in enumerate(marks):
            print(f"  chunk {i} stream {i&1}: start {e0.elapsed_time(a)*1e3:6.0f}  sweep_done {e0.elapsed_time(b_)*1e3:6.0f}  end {e0.elapsed_time(c_)*1e3:6.0f}")
for i in range(4): run(False)
run(True); run(True)
