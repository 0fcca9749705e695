"""Time spq_softmax_loss_grad16 at the LM-head shape:  python tools/loss_bench.py [M] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from llm_qat_on_gpt2_b200 import _lib
M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
V = 50257
ld = (V + 31) // 32 * 32
sb = torch.randn(M, ld, device="cuda"); tb = sb + torch.randn(M, ld, device="cuda") * 0.5
s, t = sb[:, :V], tb[:, :V]            # row-padded views, as the LM head returns them
tg = torch.randint(0, V, (M,), device="cuda")
for kind, kw in (("kl", dict(t2d=t, temperature=3.0, seq_len=256)), ("ce", dict(targets=tg))):
    for _ in range(3):
        _lib.softmax_loss_grad16(s, kind, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        _lib.softmax_loss_grad16(s, kind, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = M * V * ((8 if kind == "kl" else 4) + 2)
    print(f"{kind} M={M}: {ms*1e3:.0f} us  {nbytes/ms/1e9:.2f} TB/s algorithmic ({os.environ.get('SPQ_LOSS_THREADS','1024')} threads)")
