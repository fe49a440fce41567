"""Timeline of the mlp.3 GEMM with A-operand converters (tuning aid; run on the GPU box).
    python tools/aconv_timeline.py [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onepose_b200 import _lib
lib = _lib.load()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
resid_k = int(sys.argv[2]) if len(sys.argv) > 2 else 1
torch.manual_seed(0)
a = torch.randn(rows, 512, device="cuda")
w = torch.randn(256, 512, device="cuda") / 512 ** 0.5
x = torch.randn(rows, 256, device="cuda")
mu = torch.randn(512, device="cuda") * 0.1
rstd = torch.rand(512, device="cuda") + 0.5
bias = torch.randn(256, device="cuda") * 0.1
wp = [torch.empty(256, 512, dtype=torch.float16, device="cuda") for _ in range(2)]
xp = [torch.empty(rows, 256, dtype=torch.float16, device="cuda") for _ in range(2)]
lib.opb_debug_split(w.data_ptr(), wp[0].data_ptr(), wp[1].data_ptr(), w.numel(), None)
eye = torch.eye(256, device="cuda")
ep = [torch.empty(256, 256, dtype=torch.float16, device="cuda") for _ in range(2)]
lib.opb_debug_split(eye.data_ptr(), ep[0].data_ptr(), ep[1].data_ptr(), eye.numel(), None)
n_ctas = 148
tl = torch.zeros(n_ctas, 64, dtype=torch.int64, device="cuda")
for it in range(3):
    lib.opb_debug_split(x.data_ptr(), xp[0].data_ptr(), xp[1].data_ptr(), x.numel(), None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.opb_debug_gemm_aconv(a.data_ptr(), wp[0].data_ptr(), wp[1].data_ptr(), xp[0].data_ptr(), xp[1].data_ptr(), mu.data_ptr(), rstd.data_ptr(),
                                  bias.data_ptr(), ep[0].data_ptr() if resid_k else None, ep[1].data_ptr() if resid_k else None, rows, tl.data_ptr(), None)
    e1.record(); torch.cuda.synchronize()
    assert rc == 0, rc
ms = e0.elapsed_time(e1)
print(f"aconv mlp.3 {rows}x256x512: {ms*1e3:.1f} us, {2*rows*256*512/ms/1e9:.1f} TFLOP/s algorithmic; HBM {(rows*(2048+1024+1024))/ms/1e6:.0f} GB/s")
hn = torch.relu((a - mu) * rstd)
ref = x.double() + hn.double() @ w.double().T + bias.double()
got = (xp[0].float() + xp[1].float()) / 64.0
print("max abs err vs fp64:", float((got.double() - ref).abs().max()), "ref max", float(ref.abs().max()))
t = tl.cpu().numpy()
med = lambda v: int(np.median(v))
print("total cycles per CTA:", med(t[:, 2] - t[:, 0]))
print("converter (tile 1) raw-ready gaps :", [med(t[:, 3 + 2 * (k + 1)] - t[:, 3 + 2 * k]) for k in range(7)])
print("converter (tile 1) convert time   :", [med(t[:, 4 + 2 * k] - t[:, 3 + 2 * k]) for k in range(8)])
print("converter kb=3 (tile 1): raw-ready -> stores issued -> fence done -> arrived:", med(t[:, 56] - t[:, 9]), med(t[:, 57] - t[:, 56]), med(t[:, 10] - t[:, 57]))
print("RESID epilogue chunk 2 of tile 1: ld+stage_wait+bar | x sts, next ldg, tmem wait | compute+sts | fence :", med(t[:, 59] - t[:, 58]), med(t[:, 60] - t[:, 59]), med(t[:, 61] - t[:, 60]), med(t[:, 62] - t[:, 61]))
lead = t[::2]
print("MMA data-ready gaps (tile 1, leader):", [med(lead[:, 21 + k] - lead[:, 20 + k]) for k in range(7)])
print("MMA ready - converter done (leader, tile 1):", [med(lead[:, 20 + k] - lead[:, 4 + 2 * k]) for k in range(8)])
print("accum-ready interval per tile (leader):", [med(lead[:, 40 + 2 * (i + 1)] - lead[:, 40 + 2 * i]) for i in range(6)])
print("epilogue duration per tile            :", [med(t[:, 41 + 2 * i] - t[:, 40 + 2 * i]) for i in range(6)])
