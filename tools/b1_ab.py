"""B=1 per-frame latency through forward(data) with PDL on / off, interleaved (reference calling convention, inference.py:146)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from onepose_b200 import GATsSuperGlue, _lib, synthetic  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
hp = dict(synthetic.DEFAULT_HPARAMS)
sd = synthetic.make_state_dict(0)
model = GATsSuperGlue(hp).eval()
model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
model = model.to(dev)
N, M = 1024, 7000
db, leaves = synthetic.make_object(0, M, 8)
d3, d2 = torch.from_numpy(db)[None].to(dev), torch.from_numpy(leaves)[None].to(dev)
k2, k3 = torch.zeros(1, N, 2, device=dev), torch.zeros(1, M, 3, device=dev)
qs = [torch.from_numpy(synthetic.make_frame(f, db, N)[0])[None].to(dev) for f in range(8)]
mk = lambda q: {"keypoints2d": k2, "keypoints3d": k3, "descriptors2d_query": q, "descriptors3d_db": d3, "descriptors2d_db": d2}  # noqa: E731
for f in range(10):
    model(mk(qs[f % 8]))
torch.cuda.synchronize()
res = {0: [], 1: []}
for rep in range(12):
    for v in (0, 1):
        lib.opb_debug_set_pdl(v)
        model(mk(qs[0]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in range(20):
            model(mk(qs[f % 8]))
        e1.record()
        torch.cuda.synchronize()
        res[v].append((e0.elapsed_time(e1) / 20, 1e3 * (time.perf_counter() - t0) / 20))
lib.opb_debug_set_pdl(1)
for v in (0, 1):
    a = np.array(res[v])
    print(f"B=1 forward(data): PDL {'on ' if v else 'off'}  device {a[:, 0].mean():.4f} +- {a[:, 0].std():.4f} ms/frame   wall {a[:, 1].mean():.4f} ms/frame")
