"""A/B of the convolution operand staging (opb_debug_set_conv_halo): correctness of the encoder layers against the oracle and
time of one SuperPoint batch, one process per mode (a wrong descriptor mode must not take the others down).
    python tools/conv_halo_ab.py            # runs modes 0 (nine boxes) and 1 (halo boxes) in subprocesses
    python tools/conv_halo_ab.py MODE       # one mode"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) == 1:
    for mode in (0, 1):
        r = subprocess.run([sys.executable, __file__, str(mode)], capture_output=True, text=True, timeout=300)
        print(f"== mode {mode} (rc {r.returncode})\n{r.stdout[-1500:]}{r.stderr[-600:] if r.returncode else ''}", flush=True)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from onepose_b200 import SuperPoint, _lib, synthetic  # noqa: E402
from oracle import superpoint_oracle as O  # noqa: E402
from tests import sp_emulation as E  # noqa: E402

mode = int(sys.argv[1])
lib = _lib.load()
assert lib.opb_debug_set_conv_halo(mode) == 0
sd = synthetic.make_superpoint_state_dict(0, 4.0)
sp = SuperPoint(synthetic.SUPERPOINT_CONF).eval()
sp.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()})
sp = sp.cuda()
H, W, B = 64, 80, 2
x = torch.from_numpy(np.stack([synthetic.make_image(i, H, W) for i in (1, 2)], 0))
p = O.params_from_numpy(sd)
sp.forward_padded(x.cuda())
chans = [64, 64, 64, 64, 64, 64, 128, 128, 128, 128, 128]
stages = [0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3]
errs = []
for i in (1, 3, 4, 6, 7, 9, 10):
    sp.debug_stop_after(i)
    sp.forward_padded(x.cuda())
    h, w, P = E.stage(H, W, stages[i])
    rows = sp.debug_read(4, B * P * chans[i]).cpu().numpy().reshape(B * P, chans[i])
    errs.append(float(np.abs(rows - E.to_grid(O.encoder(p, x, upto=i).numpy())).max()))
sp.debug_stop_after(-1)
print("encoder errors (conv1b 2a 2b 3a 3b 4a 4b):", " ".join(f"{e:.1e}" for e in errs))
sp.forward_padded(x.cuda())
feat = O.encoder(p, x)
h3, w3, P3 = E.stage(H, W, 3)
lg = sp.debug_read(2, B * P3 * 128).cpu().numpy().reshape(B * P3, 128)[:, :65]
dd = sp.debug_read(3, B * P3 * 256).cpu().numpy().reshape(B * P3, 256)
sc = sp.debug_read(0, B * H * W).cpu().reshape(B, H, W)
print("logits err %.2e  dense-descriptor err %.2e  dense-score err %.2e" % (
    np.abs(E.from_grid(lg, B, 65, h3, w3) - O._conv(p, "convPb", O._conv(p, "convPa", feat), relu=False).numpy()).max(),
    np.abs(E.from_grid(dd, B, 256, h3, w3) - O._conv(p, "convDb", O._conv(p, "convDa", feat), relu=False).numpy()).max(),
    float((sc - O.dense_scores(p, feat)).abs().max())))
img = torch.from_numpy(np.stack([synthetic.make_image(100 + i, 512, 512) for i in range(8)], 0)).cuda()
for _ in range(3):
    sp.forward_padded(img)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
ev[0].record()
for i in range(6):
    sp.forward_padded(img)
    ev[i + 1].record()
torch.cuda.synchronize()
print("ms per batch:", " ".join(f"{ev[i].elapsed_time(ev[i + 1]):.3f}" for i in range(6)))
sp.set_profiling(True)
sp.forward_padded(img)
print("  ".join(f"{n} {ms:.3f}" for n, ms, _ in sp.get_profile() if n.startswith("conv")))
