"""Per-launch-kind in-stream profile of ONE frame (B=1) through match_frames at the metric shape: where the reference's own
calling convention (inference.py:146) spends its ~0.9 ms.   OPB_PROFILE_DUMP=1 python tools/b1_profile.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["OPB_PROFILE_DUMP"] = "1"
from onepose_b200 import GATsSuperGlue, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
model = GATsSuperGlue(dict(synthetic.DEFAULT_HPARAMS)).eval()
model.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in synthetic.make_state_dict(0).items()})
model = model.to(dev)
N, M = 1024, 7000
db, leaves = synthetic.make_object(0, M, 8)
model.set_object(torch.from_numpy(db).to(dev), torch.from_numpy(leaves).to(dev), reserve=(1, N))
qs = [torch.from_numpy(synthetic.make_frame(f, db, N)[0])[None].to(dev) for f in range(4)]
for f in range(6):
    model.match_frames(qs[f % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for f in range(20):
    model.match_frames(qs[f % 4])
e1.record()
torch.cuda.synchronize()
print(f"B=1 match_frames: {e0.elapsed_time(e1) / 20:.4f} ms/frame (no profiling), {model.launch_count()} launches", file=sys.stderr)
model.set_profiling(True)
model.match_frames(qs[0])
model.get_profile()          # prints the table (OPB_PROFILE_DUMP)
