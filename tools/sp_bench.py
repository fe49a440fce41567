"""The SuperPoint leg of bench.py on its own (tuning aid; also the command the ncu captures of the convolution kernels run):
    python tools/sp_bench.py [--batch 8] [--size 512] [--once]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from onepose_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--once", action="store_true", help="one warm-up and one batch only (for ncu)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
if a.once:
    import numpy as np
    from onepose_b200 import SuperPoint
    sp = SuperPoint(synthetic.SUPERPOINT_CONF).eval()
    sp.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in synthetic.make_superpoint_state_dict(0, 4.0).items()})
    sp = sp.to(dev)
    img = torch.from_numpy(np.stack([synthetic.make_image(100 + i, a.size, a.size) for i in range(a.batch)], 0)).to(dev)
    for _ in range(2):
        out = sp.forward_padded(img)
    torch.cuda.synchronize()
    print(out["counts"].tolist())
else:
    print(json.dumps({"superpoint": bench.leg_superpoint(dev, synthetic, 0, B=a.batch, H=a.size, W=a.size),
                      "pipeline": bench.leg_pipeline(dev, synthetic, B=a.batch, H=a.size, W=a.size)}))
