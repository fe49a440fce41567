"""Precision ladder (design tool, CPU): which tensor-core operand format holds the
1e-4 conf / bit-exact-index contract?

Emulates the PLANNED device dataflow (point-major activations, head-contiguous
channels, merge folded into mlp.0, Q normalised by K-mean) with every GEMM's
operands rounded the way a given tcgen05 scheme would round them, products and
sums carried in fp64 (i.e. an optimistic model of fp32 TMEM accumulation), and
compares with the fp64 oracle.

    python tools/precision_ladder.py [N M]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from onepose_b200 import synthetic  # noqa: E402
from oracle import gats_spg_oracle as oracle  # noqa: E402

H, DH = 4, 64


def split(x, fmt, scale_lo):
    """x (fp32) -> (hi, lo) as fp64 tensors holding exactly the values the
    tensor core would see; lo carries the extra 2^scale_lo factor."""
    if fmt == "tf32":
        xi = x.view(torch.int32)
        hi = ((xi + 0x1000) & ~0x1FFF).view(torch.float32)  # round-to-nearest on 13 dropped bits
        r = (x - hi) * (2.0 ** scale_lo)
        lo = (r.view(torch.int32) & ~0x1FFF).view(torch.float32)  # tensor core truncates
        return hi.double(), lo.double()
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[fmt]
    hi = x.to(dt).float()
    lo = ((x - hi) * (2.0 ** scale_lo)).to(dt).float()
    return hi.double(), lo.double()


class MM:
    """A @ B^T under an operand-rounding scheme."""

    def __init__(self, mode):
        self.mode = mode

    def __call__(self, A, Bt):
        m = self.mode
        if m == "fp32":
            return A @ Bt.T
        if m == "fp64ops":
            return (A.double() @ Bt.double().T).float()
        fmt, passes = m.split("x")
        scale = 11 if fmt == "fp16s" else 0
        fmt = fmt.rstrip("s")
        ah, al = split(A, fmt, scale)
        bh, bl = split(Bt, fmt, scale)
        out = ah @ bh.T
        if passes == "3":
            out = out + (ah @ bl.T + al @ bh.T) * (2.0 ** -scale)
        elif passes == "2":     # activations single, weights split (or vice versa): hi*hi + hi*lo
            out = out + (ah @ bl.T) * (2.0 ** -scale)
        elif passes == "4":
            out = out + (ah @ bl.T + al @ bh.T) * (2.0 ** -scale) + (al @ bl.T) * (2.0 ** (-2 * scale))
        return out.float()


def pack(sd):
    """Host-side weight pre-pack (what load_weights will do), fp64 folding then fp32."""
    P = {}
    # head-contiguous channel permutation: new c' = h*64 + d  <- old c = d*4 + h (GATs_SuperGlue.py:97)
    perm = np.array([d * H + h for h in range(H) for d in range(DH)])
    for i, name in enumerate(synthetic.GNN_LAYERS):
        p = f"gnn.layers.{i}"
        if name == "GATs":
            W = sd[f"{p}.W"].astype(np.float64)
            a = sd[f"{p}.a"].astype(np.float64)[:, 0]
            P[i] = dict(wa2=torch.tensor(W @ a[:256], dtype=torch.float32),
                        wa3=torch.tensor(W @ a[256:], dtype=torch.float32))
        else:
            g = lambda k: sd[f"{p}.{k}"].astype(np.float64)
            Wq, Wk, Wv = (g(f"attn.proj.{j}.weight")[:, :, 0][perm] for j in range(3))
            bq, bk, bv = (g(f"attn.proj.{j}.bias")[perm] for j in range(3))
            Wm = g("attn.merge.weight")[:, :, 0][:, perm]     # input channels permuted
            bm = g("attn.merge.bias")
            W0 = g("mlp.0.weight")[:, :, 0]
            b0 = g("mlp.0.bias")
            W0a, W0b = W0[:, :256], W0[:, 256:]
            f32 = lambda x: torch.tensor(x, dtype=torch.float32)
            P[i] = dict(Wqkv=f32(np.concatenate([Wq, Wk, Wv], 0)), bqkv=f32(np.concatenate([bq, bk, bv])),
                        Wm=f32(Wm), bm=f32(bm), W0=f32(W0), b0=f32(b0),
                        W0fold=f32(np.concatenate([W0a, W0b @ Wm], 1)), b0fold=f32(W0b @ bm + b0),
                        W1=f32(g("mlp.3.weight")[:, :, 0]), b1=f32(g("mlp.3.bias")))
    P["Wf"] = torch.tensor(sd["final_proj.weight"][:, :, 0])
    P["bf"] = torch.tensor(sd["final_proj.bias"])
    return P


def elu1(x):
    return torch.where(x > 0, x + 1, torch.exp(x))


def side_qkv(x, L, mm):
    qkv = mm(x, L["Wqkv"]) + L["bqkv"]
    n = x.shape[0]
    Q = elu1(qkv[:, :256]).reshape(n, H, DH)
    K = elu1(qkv[:, 256:512]).reshape(n, H, DH)
    V = qkv[:, 512:].reshape(n, H, DH)
    Kmean = K.mean(0)                                             # [H, d]
    KVmean = torch.einsum("nhd,nhq->hdq", K.double(), V.double()).float() / n
    return Q, Kmean, KVmean


def side_update(x, Q, Kmean, KVmean, m_src, L, mm, fold):
    n = x.shape[0]
    Z = 1.0 / ((Q * Kmean[None]).sum(-1, keepdim=True) + 1e-6 / m_src)
    Qn = Q * Z
    if fold == "G":     # dynamic per-frame weight G = blockdiag(KVmean) @ W0m^T
        W0m = L["W0fold"][:, 256:].reshape(512, H, DH)            # [c, h, q]
        G = torch.einsum("hdq,chq->chd", KVmean.double(), W0m.double()).float().reshape(512, 256)
        hid = mm(torch.cat([x, Qn.reshape(n, 256)], 1), torch.cat([L["W0fold"][:, :256], G], 1)) + L["b0fold"]
    else:
        msg = torch.einsum("nhd,hdq->nhq", Qn, KVmean).reshape(n, 256)
        if fold == "merge":
            hid = mm(torch.cat([x, msg], 1), L["W0fold"]) + L["b0fold"]
        else:
            message = mm(msg, L["Wm"]) + L["bm"]
            hid = mm(torch.cat([x, message], 1), L["W0"]) + L["b0"]
    mu = hid.mean(0, keepdim=True)
    var = hid.var(0, unbiased=False, keepdim=True)
    hn = torch.relu((hid - mu) * torch.rsqrt(var + 1e-5))
    return x + mm(hn, L["W1"]) + L["b1"]


def planned_forward(P, q, db, leaves, mm, mm_score, fold="merge", Lf=8):
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
            Qq, Kq, KVq = side_qkv(xq, L, mm)
            Qd, Kd, KVd = side_qkv(xd, L, mm)
            if name == "self":
                nq = side_update(xq, Qq, Kq, KVq, xq.shape[0], L, mm, fold)
                nd = side_update(xd, Qd, Kd, KVd, xd.shape[0], L, mm, fold)
            else:
                nq = side_update(xq, Qq, Kd, KVd, xd.shape[0], L, mm, fold)
                nd = side_update(xd, Qd, Kq, KVq, xq.shape[0], L, mm, fold)
            xq, xd = nq, nd
    pq = mm(xq, P["Wf"]) + P["bf"]
    pd = mm(xd, P["Wf"]) + P["bf"]
    pq = pq / pq.norm(dim=1, keepdim=True).clamp_min(1e-12)
    pd = pd / pd.norm(dim=1, keepdim=True).clamp_min(1e-12)
    s = mm_score(pq, pd) / 0.07
    return torch.softmax(s, 0) * torch.softmax(s, 1)


def main():
    N, M = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (512, 2048)
    torch.set_num_threads(8)
    sd = synthetic.make_state_dict(0, damped=True)
    data = synthetic.make_batch(5, [51], N, M, 8)
    hp = synthetic.DEFAULT_HPARAMS
    ref64 = oracle.forward(oracle.params_from_numpy(sd, torch.float64), data, hp, dtype=torch.float64)
    ref32 = oracle.forward(oracle.params_from_numpy(sd), data, hp)
    c64 = ref64["conf_matrix"][0]
    c32 = ref32["conf_matrix"][0]
    top2 = c64.topk(2, dim=1).values
    print(f"N={N} M={M}: ref fp32 vs fp64 max|dconf|={float((c32.double() - c64).abs().max()):.2e}; "
          f"min row top1-top2 gap (fp64) = {float((top2[:, 0] - top2[:, 1]).min()):.2e}")
    P = pack(sd)
    q, db, lv = (torch.tensor(data[k][0]) for k in ("descriptors2d_query", "descriptors3d_db", "descriptors2d_db"))
    print(f"{'gemm':>9} {'score':>9} {'fold':>6} | {'max|dconf| vs fp64':>18} {'vs ref fp32':>12} | row-flips col-flips (vs ref fp32)")
    for mode, smode, fold in [("fp32", "fp32", "none"), ("fp32", "fp32", "merge"), ("fp32", "fp32", "G"),
                              ("bf16x1", "fp32", "merge"), ("bf16x3", "bf16x3", "merge"), ("fp16x3", "fp16x3", "merge"),
                              ("fp16sx3", "fp16sx3", "merge"), ("tf32x3", "tf32x3", "merge"),
                              ("fp16sx3", "fp16sx3", "G"), ("fp16x3", "fp16x3", "G"),
                              ("fp16x2", "fp16x3", "merge"), ("tf32x1", "tf32x3", "merge"), ("fp16x1", "fp16x3", "merge"),
                              ("fp16sx4", "fp16sx4", "merge")]:
        t0 = time.time()
        c = planned_forward(P, q, db, lv, MM(mode), MM(smode), fold)
        d64 = float((c.double() - c64).abs().max())
        d32 = float((c - c32).abs().max())
        rf = int((c.argmax(1) != c32.argmax(1)).sum())
        cf = int((c.argmax(0) != c32.argmax(0)).sum())
        print(f"{mode:>9} {smode:>9} {fold:>6} | {d64:18.2e} {d32:12.2e} | {rf:5d} {cf:5d}   ({time.time() - t0:.1f}s)")


if __name__ == "__main__":
    main()
