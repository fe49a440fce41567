"""Per-CTA timeline of the tcgen05 GEMM core (tuning aid; run on the GPU box).
    python tools/gemm_timeline.py rows n_out K"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from onepose_b200 import _lib
lib = _lib.load()
rows, n_out, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (32768, 512, 512)
dbg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
a = torch.randn(rows, K, device="cuda"); b = torch.randn(n_out, K, device="cuda") / K ** 0.5
pl = [torch.empty(rows, K, dtype=torch.float16, device="cuda") for _ in range(2)] + [torch.empty(n_out, K, dtype=torch.float16, device="cuda") for _ in range(2)]
lib.opb_debug_split(a.data_ptr(), pl[0].data_ptr(), pl[1].data_ptr(), a.numel(), None)
lib.opb_debug_split(b.data_ptr(), pl[2].data_ptr(), pl[3].data_ptr(), b.numel(), None)
c = torch.empty(rows, n_out, device="cuda")
n_tiles = (rows // 128) * (n_out // 256)
n_ctas = min(n_tiles, 148)
tl = torch.zeros(n_ctas, 64, dtype=torch.int64, device="cuda")
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = lib.opb_debug_gemm_timeline(*(p.data_ptr() for p in pl), c.data_ptr(), rows, n_out, K, tl.data_ptr(), dbg, None)
    e1.record(); torch.cuda.synchronize()
    assert rc == 0
ms = e0.elapsed_time(e1)
print(f"GEMM {rows}x{n_out}x{K}: {ms*1e3:.1f} us, {2*rows*n_out*K/ms/1e9:.1f} TFLOP/s algorithmic ({6*rows*n_out*K/ms/1e9:.0f} executed), {n_tiles} tiles on {n_ctas} CTAs")
print("dbg variant", dbg)
if dbg != 9:
    ref = (a.double() @ b.double().T).float()
    print("max rel err vs fp64:", float((c - ref).abs().max() / ref.abs().max()))
t = tl.cpu().numpy()
nkb = K // 64
d = lambda i, j: np.median(t[:, i] - t[:, j])
ntl = min(8, n_tiles // n_ctas)
print(f"median cycles per CTA: total {d(2,0):.0f} | setup {d(1,0):.0f} | setup->first data {d(20,1):.0f} | first tile mainloop {d(40,20):.0f}")
print("data-ready gaps (tile 0):", [int(np.median(t[:, 20+k+1]-t[:, 20+k])) for k in range(min(nkb,16)-1)])
print("per tile: accum-ready interval:", [int(d(40+2*(i+1), 40+2*i)) for i in range(ntl-1)])
print("per tile: epilogue duration   :", [int(d(41+2*i, 40+2*i)) for i in range(ntl)])
if dbg == 9:
    print("EPI_QKV chunk (tile 2, 2nd chunk of group 0): tmem ld+wait | stage_wait+bar | compute+sts | fence+bar+store:", int(d(4,3)), int(d(5,4)), int(d(6,5)), int(d(7,6)))
print("epilogue chunk phases (tile 1; ld, wait_read+bar, stage+fence+bar, next):", [[int(d(4+4*c,3+4*c)), int(d(5+4*c,4+4*c)), int(d(6+4*c,5+4*c)), int(d(7+4*c,6+4*c)) if c<3 else -1] for c in range(4)])
