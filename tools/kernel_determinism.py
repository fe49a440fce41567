"""Debug aid (GPU box): are the fp16 state kernel and the EPI_QKV GEMM bit-reproducible on identical inputs?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onepose_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
B, N, M = 8, 1024, 7000
R = 1024 + 7168
rows = B * R
kv = (torch.randn(rows, 512, device="cuda") * 64).half()
kv.view(B, R, 512)[:, 1024 + M:, :] = 0
kv[:, :256] = kv[:, :256].abs()
outs = []
for it in range(4):
    part = torch.full((rows // 256, 4, 4160), float("nan"), device="cuda")
    assert lib.opb_debug_kv_state_h(kv.data_ptr(), B, N, M, part.data_ptr(), None) == 0
    torch.cuda.synchronize()
    outs.append(part)
print("kv_state_h: identical", [torch.equal(outs[0], o) for o in outs[1:]], " any nan", bool(torch.isnan(outs[0]).any()),
      " max diff", max(float((outs[0] - o).abs().max()) for o in outs[1:]))
ref = torch.einsum("srhd,srhq->shdq", kv[:, :256].float().view(rows // 256, 256, 4, 64), kv[:, 256:].float().view(rows // 256, 256, 4, 64)) / 4096
got = outs[0][:, :, :4096].view(rows // 256, 4, 64, 64)
print("kv_state_h vs torch: max rel err", float((got - ref).abs().max() / ref.abs().max()))
ks = kv[:, :256].float().view(rows // 256, 256, 4, 64).sum(1) / 64
print("ksum vs torch: max rel err", float((outs[0][:, :, 4096:] - ks).abs().max() / ks.abs().max()))
# EPI_QKV GEMM (k,v projection form)
a = torch.randn(rows, 256, device="cuda"); b = torch.randn(512, 256, device="cuda") / 16
pl = [torch.empty(rows, 256, dtype=torch.float16, device="cuda") for _ in range(2)] + [torch.empty(512, 256, dtype=torch.float16, device="cuda") for _ in range(2)]
lib.opb_debug_split(a.data_ptr(), pl[0].data_ptr(), pl[1].data_ptr(), a.numel(), None)
lib.opb_debug_split(b.data_ptr(), pl[2].data_ptr(), pl[3].data_ptr(), b.numel(), None)
res = []
for it in range(4):
    c = torch.zeros(rows, 512, dtype=torch.float16, device="cuda")
    assert lib.opb_debug_gemm_timeline(*(p.data_ptr() for p in pl), c.data_ptr(), rows, 512, 256, None, 9, None) == 0
    torch.cuda.synchronize()
    res.append(c)
print("EPI_QKV gemm: identical", [torch.equal(res[0], r) for r in res[1:]])
