"""Per-GEMM-site precision ladder under a weight-damping sweep (design tool, CPU).

Question: the device runs every GEMM as three fp16 passes
(A_hi.B_hi + A_hi.B_lo + A_lo.B_hi).  Which SITES need all three, and in which weight regime?

Emulates the device dataflow of onepose_b200/csrc (fp16 hi/lo planes with the 2^6 pre-scale, single fp16 plane for the
linear-attention state operands, G fold, InstanceNorm from fp32 hidden, identity-block residual) with products and sums
carried in fp64 (an optimistic model of the fp32 TMEM accumulator), one site at a time degraded to fewer passes, and
compares with the fp64 referee (oracle in float64) -- next to the reference's own fp32 forward.

    python tools/precision_sites.py [N M] [--scales 0.02,0.1,0.3,1.0] [--out profiles/xyz.json]

Sites: kv_proj, q_proj, mlp0 (x and Q' halves), gfold, mlp3, final, score.
Modes: x3 (product), x2w = hi.hi + hi.lo(weights)  [activation lo dropped], x2a = hi.hi + lo(act).hi [weight lo dropped], x1.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from onepose_b200 import synthetic  # noqa: E402
from oracle import gats_spg_oracle as oracle  # noqa: E402
from precision_ladder import DH, H, elu1, pack  # noqa: E402

SITES = ["kv_proj", "q_proj", "mlp0", "gfold", "mlp3", "final", "score"]


def planes(x):
    """fp32 -> the two fp16 planes of the device format, as float64 holding the exact plane values / 64."""
    xs = x.float() * 64.0
    hi = xs.to(torch.float16).float()
    lo = (xs - hi).to(torch.float16).float()
    return hi.double() / 64.0, lo.double() / 64.0


def mm(A, Bt, mode):
    """A @ Bt^T with A = activations, Bt = weights ([n_out, K])."""
    ah, al = planes(A)
    bh, bl = planes(Bt)
    out = ah @ bh.T
    if mode in ("x3", "x2w"):
        out = out + ah @ bl.T
    if mode in ("x3", "x2a"):
        out = out + al @ bh.T
    return out.float()


def fp16_plane(x):
    return (x.float() * 64.0).to(torch.float16).double() / 64.0


def side_state(x, L, modes):
    kv = mm(x, L["Wqkv"][256:], modes["kv_proj"]) + L["bqkv"][256:]
    n = x.shape[0]
    K = fp16_plane(elu1(kv[:, :256])).reshape(n, H, DH)          # single fp16 plane, one tensor pass
    V = fp16_plane(kv[:, 256:]).reshape(n, H, DH)
    return K.mean(0).float(), (torch.einsum("nhd,nhq->hdq", K, V) / n).float()


def side_update(x, Kmean, KVmean, m_src, L, modes):
    n = x.shape[0]
    q = mm(x, L["Wqkv"][:256], modes["q_proj"]) + L["bqkv"][:256]
    Q = elu1(q).reshape(n, H, DH)
    Qn = (Q / ((Q * Kmean[None]).sum(-1, keepdim=True) + 1e-6 / m_src)).reshape(n, 256)
    W0m = L["W0fold"][:, 256:]                                        # [512, 256] head-contiguous
    bd = torch.zeros(256, 256)
    for h in range(H):
        bd[h * DH:(h + 1) * DH, h * DH:(h + 1) * DH] = KVmean[h]     # [d, q] block
    G = mm(W0m, bd, modes["gfold"])                                   # G[c, h*64+d] = sum_q W0m[c, h*64+q] KV[h][d][q]
    hid = mm(torch.cat([x, Qn], 1), torch.cat([L["W0fold"][:, :256], G], 1), modes["mlp0"]) + L["b0fold"]
    mu = hid.double().mean(0, keepdim=True)
    var = (hid.double() ** 2).mean(0, keepdim=True) - mu * mu
    hn = torch.relu((hid - mu.float()) * torch.rsqrt(var.clamp_min(0) + 1e-5).float())
    delta = mm(hn, L["W1"], modes["mlp3"])
    xh, xl = planes(x)
    return (delta.double() + xh + xl).float() + L["b1"]               # identity K-block: the residual is exact in the accumulator


def device_forward(P, q, db, leaves, modes, Lf=8):
    xq, xd = q.T.contiguous(), db.T.contiguous()
    lv = leaves.T.contiguous()
    M = xd.shape[0]
    for i, name in enumerate(synthetic.GNN_LAYERS):
        L = P[i]
        if name == "GATs":
            s2 = (lv @ L["wa2"]).reshape(M, Lf)
            s3 = xd @ L["wa3"]
            e = torch.cat([2 * s3[:, None], s3[:, None] + s2], 1)
            e = torch.where(e > 0, e, 0.2 * e)
            att = torch.softmax(e, 1)
            hp = att[:, :1] * xd + (att[:, 1:, None] * lv.reshape(M, Lf, 256)).sum(1)
            xd = torch.where(hp > 0, hp, torch.expm1(hp))
        else:
            Kq, KVq = side_state(xq, L, modes)
            Kd, KVd = side_state(xd, L, modes)
            if name == "self":
                nq = side_update(xq, Kq, KVq, xq.shape[0], L, modes)
                nd = side_update(xd, Kd, KVd, xd.shape[0], L, modes)
            else:
                nq = side_update(xq, Kd, KVd, xd.shape[0], L, modes)
                nd = side_update(xd, Kq, KVq, xq.shape[0], L, modes)
            xq, xd = nq, nd
    pq = mm(xq, P["Wf"], modes["final"]) + P["bf"]
    pd = mm(xd, P["Wf"], modes["final"]) + P["bf"]
    pq = pq / pq.norm(dim=1, keepdim=True).clamp_min(1e-12)
    pd = pd / pd.norm(dim=1, keepdim=True).clamp_min(1e-12)
    s = mm(pq, pd, modes["score"]) / 0.07
    return torch.softmax(s, 0) * torch.softmax(s, 1)


def metrics(c, c64):
    d = float((c.double() - c64).abs().max())
    top2 = c64.topk(2, dim=1).values
    decided = (top2[:, 0] - top2[:, 1]) > 1e-5 * top2[:, 0]
    flips = int((c.argmax(1) != c64.argmax(1))[decided].sum())
    top2c = c64.topk(2, dim=0).values
    decided_c = (top2c[0] - top2c[1]) > 1e-5 * top2c[0]
    flips_c = int((c.argmax(0) != c64.argmax(0))[decided_c].sum())
    return d, flips, flips_c, d / float(c64.max())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shape", nargs="*", type=int, default=[512, 2048])
    ap.add_argument("--scales", default="0.02,0.1,0.3,1.0")
    ap.add_argument("--modes", default="x2w,x2a,x1")
    ap.add_argument("--out", default=None)
    ap.add_argument("--threads", type=int, default=8)
    args = ap.parse_args()
    N, M = args.shape
    torch.set_num_threads(args.threads)
    hp = synthetic.DEFAULT_HPARAMS
    data = synthetic.make_batch(5, [51], N, M, 8)
    q, db, lv = (torch.tensor(data[k][0]) for k in ("descriptors2d_query", "descriptors3d_db", "descriptors2d_db"))
    report = {"N": N, "M": M, "scales": [], "note": "max|dconf| vs the fp64 referee; flips = raw arg-max differences vs fp64 where fp64's "
              "top-1/top-2 gap exceeds 1e-5 relative (rows, columns); emulation carries sums in fp64"}
    for scale in [float(s) for s in args.scales.split(",")]:
        sd = synthetic.make_state_dict(0, mlp3_scale=scale)
        t0 = time.time()
        c64 = oracle.forward(oracle.params_from_numpy(sd, torch.float64), data, hp, dtype=torch.float64)["conf_matrix"][0]
        c32 = oracle.forward(oracle.params_from_numpy(sd), data, hp)["conf_matrix"][0]
        top2 = c64.topk(2, dim=1).values
        entry = {"mlp3_scale": scale, "conf_max": float(c64.max()), "matches_over_0.2": int((c64.max(1).values > 0.2).sum()),
                 "min_rel_gap_top1_top2": float(((top2[:, 0] - top2[:, 1]) / top2[:, 0]).min()), "rows": []}
        d, f, fc, rel = metrics(c32, c64)
        entry["rows"].append({"config": "reference fp32 (oracle)", "max_abs_dconf": d, "rel_to_conf_max": rel, "row_flips": f, "col_flips": fc})
        print(f"\n== mlp3_scale {scale}: conf.max {entry['conf_max']:.3f}, {entry['matches_over_0.2']} rows > 0.2; "
              f"reference fp32 vs fp64: {d:.2e}, flips {f}/{fc}  ({time.time() - t0:.0f}s)")
        P = pack(sd)
        base = {s: "x3" for s in SITES}
        cases = [("all x3 (product)", base)]
        for site in SITES:
            for mode in args.modes.split(","):
                cases.append((f"{site} {mode}", dict(base, **{site: mode})))
        for name, modes in cases:
            t0 = time.time()
            c = device_forward(P, q, db, lv, modes)
            d, f, fc, rel = metrics(c, c64)
            entry["rows"].append({"config": name, "max_abs_dconf": d, "rel_to_conf_max": rel, "row_flips": f, "col_flips": fc})
            print(f"  {name:22s} max|dconf| {d:.2e} ({rel:.1e} of conf.max)  flips {f:3d}/{fc:3d}   ({time.time() - t0:.0f}s)", flush=True)
        report["scales"].append(entry)
    if args.out:
        json.dump(report, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
