"""Design check (CPU): may the linear-attention state  KV = mean_rows elu1(K)^T V  take its operands in SINGLE fp16?

Emulates the device dataflow of tools/precision_ladder.py (every GEMM in the 3-pass fp16-split scheme) and additionally
rounds K = elu1(k) and V to fp16 / bf16 before the state reduction -- what onepose_b200/csrc/kv_state_tc.cu
(kv_state_h_kernel) sees.  Reports the end-to-end error of the matching cosine (conf error ~= conf * dcos / 0.07) against
an all-fp64 run, for damped and undamped synthetic weights (undamped: attention deltas are O(1) of the features) and for
segments as short as 16 rows (the averaging argument is weakest there).

    python tools/kv_precision.py [N M]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import precision_ladder as pl  # noqa: E402
from onepose_b200 import synthetic  # noqa: E402

KV = {"mode": "exact"}


def side_qkv(x, L, mm):
    qkv = mm(x, L["Wqkv"]) + L["bqkv"]
    n = x.shape[0]
    Q = pl.elu1(qkv[:, :256]).reshape(n, pl.H, pl.DH)
    K = pl.elu1(qkv[:, 256:512]).reshape(n, pl.H, pl.DH)
    V = qkv[:, 512:].reshape(n, pl.H, pl.DH)
    Kr, Vr = K, V
    if KV["mode"] == "fp16":
        Kr, Vr = K.half().float(), V.half().float()
    elif KV["mode"] == "bf16":
        Kr, Vr = K.bfloat16().float(), V.bfloat16().float()
    Kmean = Kr.mean(0)
    KVmean = torch.einsum("nhd,nhq->hdq", Kr.double(), Vr.double()).float() / n
    return Q, Kmean, KVmean


pl.side_qkv = side_qkv


def descriptors(P, q, db, lv, mm):
    got = {}

    def grab(a, b):
        got["pq"], got["pd"] = a, b
        return a[:8] @ b[:8].T

    pl.planned_forward(P, q, db, lv, mm, grab, "G")
    return got["pq"], got["pd"]


def main():
    sizes = [(int(sys.argv[1]), int(sys.argv[2]))] if len(sys.argv) > 2 else [(16, 40), (64, 96), (512, 2048)]
    torch.set_num_threads(16)
    for N, M in sizes:
        for damped in (True, False):
            sd = synthetic.make_state_dict(0, damped=damped)
            data = synthetic.make_batch(5, [51], N, M, 8)
            P = pl.pack(sd)
            q, db, lv = (torch.tensor(data[k][0]) for k in ("descriptors2d_query", "descriptors3d_db", "descriptors2d_db"))
            KV["mode"] = "exact"
            pq0, pd0 = descriptors(P, q, db, lv, pl.MM("fp64ops"))
            cos0 = pq0.double() @ pd0.double().T
            for kvm, mmode in (("exact", "fp16sx3"), ("fp16", "fp16sx3"), ("bf16", "fp16sx3"), ("exact", "fp16x1")):
                KV["mode"] = kvm
                pq, pd = descriptors(P, q, db, lv, pl.MM(mmode))
                cos = pq.double() @ pd.double().T
                print(f"N={N:5d} M={M:5d} damped={int(damped)} state operands={kvm:6s} gemm={mmode:8s}: max|dcos| {float((cos - cos0).abs().max()):.2e}", flush=True)


if __name__ == "__main__":
    main()
