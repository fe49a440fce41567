"""Debug aid (GPU box): is frame b of a batch bit-identical to the same frame run alone?  Prints where they diverge."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from onepose_b200 import synthetic, _lib
from onepose_b200.matcher import GATsSuperGlue

hp = dict(synthetic.DEFAULT_HPARAMS)
sd = synthetic.make_state_dict(0)
B, N, M = int(os.environ.get("NB", "8")), 1024, 7000
data = synthetic.make_batch(7, list(range(70, 70 + B)), N, M, 8)
cuda = lambda d: {k: torch.from_numpy(v).cuda() for k, v in d.items()}
lib = _lib.load()

def run(d):
    m = GATsSuperGlue(dict(hp), gemm_backend=os.environ.get("BACKEND", "tcgen05")).eval()
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    m = m.cuda()
    m(cuda(d))
    out = {k: v.clone() for k, v in m.last_batched.items()}
    rows = C.c_int64(0)
    nb = d["descriptors2d_query"].shape[0]
    R = 1024 + 7168
    x = torch.empty(nb * R * 256, device="cuda")
    rc = lib.opb_debug_read(m._handle, 0, x.data_ptr(), x.numel(), C.byref(rows), None)
    torch.cuda.synchronize()
    return out, x.view(nb, R, 256) if rc == 0 else None

a, xa = run(data)
a2, xa2 = run(data)
if os.environ.get("QUICK"):
    print("batch run twice: conf equal", torch.equal(a["conf_matrix"], a2["conf_matrix"]), "x equal", torch.equal(xa, xa2), " max|dx|", float((xa - xa2).abs().max()))
    sys.exit(0)
print("batch run twice: conf equal", torch.equal(a["conf_matrix"], a2["conf_matrix"]), "x equal", None if xa is None else torch.equal(xa, xa2))
for f in (0, B - 1):
    one = {k: v[f:f + 1] for k, v in data.items()}
    s, xs = run(one)
    dc = float((s["conf_matrix"][0] - a["conf_matrix"][f]).abs().max())
    if xs is not None:
        dxq = float((xs[0, :1024] - xa[f, :1024]).abs().max())
        dxd = float((xs[0, 1024:1024 + 7000] - xa[f, 1024:1024 + 7000]).abs().max())
    else:
        dxq = dxd = float("nan")
    print(f"frame {f}: solo vs batch  max|dconf| {dc:.3e}   max|dx| query {dxq:.3e}  3d {dxd:.3e}")
pair = {k: v[3:5] for k, v in data.items()}
if B < 5: sys.exit(0)
p, xp = run(pair)
print("frames [3,4] vs batch: dconf", float((p["conf_matrix"][0] - a["conf_matrix"][3]).abs().max()), " dx", float((xp[0] - xa[3]).abs().max()))
