"""Developer tool (GPU box): event timeline of the staged two-chunk pipeline (full-length C2 batch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_chinese_e2e_b200 import _lib
from oracle.synth import make_case

B, T, V, U = 256, 400, 4234, 50
split = float(sys.argv[1]) if len(sys.argv) > 1 else 0.8
c = make_case(B, T, V, U, 1002, full_lengths=True)
x = c["logits"].cuda(); tg = c["targets"].cuda(); il = c["input_lengths"].cuda(); tl = c["target_lengths"].cuda()
L = _lib.lib()
cut = int(round(B * split / 2)) * 2
bounds = [0, cut, B]
wsb = [_lib.workspace_bytes(bounds[i+1]-bounds[i], T, V, U) for i in range(2)]
ws = [torch.empty(w, dtype=torch.uint8, device="cuda") for w in wsb]
nll = torch.empty(B, device="cuda"); sums = torch.zeros(2, 4, device="cuda"); grad = torch.empty_like(x)
side = torch.cuda.Stream(priority=-1 if os.environ.get('HIPRI','1')=='1' else 0)
ev = lambda: torch.cuda.Event(enable_timing=True)
def call(ci, stages, s):
    lo, n = bounds[ci], bounds[ci+1]-bounds[ci]
    rc = L.ctcb200_loss_grad_stages(stages, x.data_ptr()+lo*T*V*4, tg.data_ptr()+lo*U*8, U, tg.numel()-lo*U, il.data_ptr()+lo*8, tl.data_ptr()+lo*8,
                                    n, T, V, U, 0, 0, 1, 1.0/B, nll.data_ptr()+lo*4, sums.data_ptr()+ci*16, grad.data_ptr()+lo*T*V*4,
                                    ws[ci].data_ptr(), wsb[ci], s.cuda_stream)
    assert rc == 0, rc
def run(show):
    main = torch.cuda.current_stream()
    marks = {}
    e0 = ev(); e0.record(main)
    evs, evl = [ev(), ev()], [ev(), ev()]
    for ci in range(2):
        call(ci, 1, main); evs[ci].record(main); side.wait_event(evs[ci])
        a = ev(); a.record(side); call(ci, 2, side); evl[ci].record(side); marks[f"lattice{ci}"] = (a, evl[ci])
    pe = []
    for ci in range(2):
        main.wait_event(evl[ci]); a = ev(); a.record(main); call(ci, 4, main); b = ev(); b.record(main); pe.append((a, b))
    e1 = ev(); e1.record(main); torch.cuda.synchronize()
    if show:
        t = lambda e: e0.elapsed_time(e) * 1e3
        print(f"split={split} total {t(e1):.0f} us | sweep a done {t(evs[0]):.0f}, sweep b done {t(evs[1]):.0f} | "
              f"lattice a {t(marks['lattice0'][0]):.0f}-{t(marks['lattice0'][1]):.0f}, lattice b {t(marks['lattice1'][0]):.0f}-{t(marks['lattice1'][1]):.0f} | "
              f"patch a {t(pe[0][0]):.0f}-{t(pe[0][1]):.0f}, patch b {t(pe[1][0]):.0f}-{t(pe[1][1]):.0f}")
for i in range(4): run(False)
run(True); run(True)
