"""Debug aid (GPU box): run the same B-frame batch twice on fresh matchers under several feature-flag settings and report
whether the residual stream / conf are bit-identical (one process: the flags are read at opb_create)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from onepose_b200 import synthetic, _lib
from onepose_b200.matcher import GATsSuperGlue

hp = dict(synthetic.DEFAULT_HPARAMS)
sd = synthetic.make_state_dict(0)
B, N, M = int(os.environ.get("NB", "8")), 1024, 7000
data = {k: torch.from_numpy(v).cuda() for k, v in synthetic.make_batch(7, list(range(70, 70 + B)), N, M, 8).items()}
lib = _lib.load()
FLAGS = ["OPB_ACONV", "OPB_KV_HALF", "OPB_SPLIT_Q", "OPB_RESID_K", "OPB_FUSE"]

def run():
    m = GATsSuperGlue(dict(hp), gemm_backend="tcgen05").eval()
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=True)
    m = m.cuda()
    m(data)
    rows = C.c_int64(0)
    R = 1024 + 7168
    x = torch.empty(B * R * 256, device="cuda")
    lib.opb_debug_read(m._handle, 0, x.data_ptr(), x.numel(), C.byref(rows), None)
    torch.cuda.synchronize()
    return m.last_batched["conf_matrix"].clone(), x

for setting in sys.argv[1:] or ["default"]:
    for f in FLAGS:
        os.environ.pop(f, None)
    if setting != "default":
        for kv in setting.split(","):
            k, v = kv.split("=")
            os.environ[k] = v
    res = [run() for _ in range(3)]
    eq = [torch.equal(res[0][1], r[1]) for r in res[1:]]
    dx = max(float((res[0][1] - r[1]).abs().max()) for r in res[1:])
    nbad = max(int((res[0][1] != r[1]).sum()) for r in res[1:])
    print(f"{setting:50s} x identical {eq}  max|dx| {dx:.3e}  differing elements {nbad}", flush=True)
