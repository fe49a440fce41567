"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (torch CPU ops, functional, no nn.Module) of the reference's
GATsSPG 2D-3D matching forward, used ONLY by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg as the checker and the CPU baseline.  Nothing under
``onepose_b200/`` may import this file.

Parity pin: the reference ships no tests / golden vectors, so the
oracle is pinned against outputs of the reference module itself, imported from
``/root/reference`` in the build container by ``tests/golden/make_golden.py``
and committed as ``tests/golden/*.npz`` (``tests/test_oracle_golden.py``).

Each function cites the reference lines it restates.  The computation is kept
"as written" (channel-first tensors, the redundant ``h @ W`` products of the
GATs layer included) so that timing this port on host cores is a fair stand-in
for timing the reference's own CPU forward.

All functions take ``params``: a dict keyed by the reference state-dict names
(``gnn.layers.{i}...``, ``final_proj.*``) holding torch tensors of the working
dtype (fp32 = the reference's arithmetic, fp64 = accuracy referee).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

GNN_LAYERS = ["GATs", "self", "cross"] * 4  # GATs_SuperGlue.py:162
NUM_HEADS = 4                               # GATs_SuperGlue.py:43


def params_from_numpy(sd, dtype=torch.float32):
    return {k: torch.as_tensor(v).to(dtype) for k, v in sd.items()}


# --------------------------------------------------------------------------
# GATs layer -- reference GATs.py:35-88
# --------------------------------------------------------------------------
def gats_layer(h2, h3, W, a, *, alpha=0.2, include_self=True, additional=False,
               with_linear_transform=False):
    """h2: [B, M*L, D] leaves (point-major), h3: [B, M, D] -> [B, M, D]."""
    B, M, D = h3.shape
    L = h2.shape[1] // M                                   # GATs.py:38
    wh2, wh3 = h2 @ W, h3 @ W                              # GATs.py:40-41
    Do = W.shape[1]
    s2 = (wh2 @ a[:Do]).reshape(B, M, L, 1)                # GATs.py:78-79
    s3 = wh3 @ a[Do:]                                      # GATs.py:80   [B, M, 1]
    if include_self:
        s2 = torch.cat([s3[:, :, None], s2], dim=2)        # GATs.py:83-85
    e = F.leaky_relu(s3[:, :, None] + s2, alpha)           # GATs.py:87-88
    att = torch.softmax(e, dim=2)                          # GATs.py:44
    nb_h = h2.reshape(B, M, L, D)
    nb_wh = wh2.reshape(B, M, L, Do)
    if include_self:
        nb_h = torch.cat([h3[:, :, None], nb_h], dim=2)    # GATs.py:52-54
        nb_wh = torch.cat([wh3[:, :, None], nb_wh], dim=2)  # GATs.py:49-51
        src = nb_wh if with_linear_transform else nb_h
        hp = (att * src).sum(dim=2)                        # GATs.py:56-59
        if additional:
            hp = hp + h3                                   # GATs.py:61-62
    else:
        src = nb_wh if with_linear_transform else nb_h
        hp = (att * src).sum(dim=2) / 2.0 + (wh3 if with_linear_transform else h3)  # GATs.py:64-67
    return F.elu(hp)                                       # GATs.py:69-70 (concat=True)


# --------------------------------------------------------------------------
# AttentionPropagation -- reference GATs_SuperGlue.py:69-128
# --------------------------------------------------------------------------
def _conv1x1(x, w, b):
    """nn.Conv1d(kernel_size=1): x [B, Cin, n], w [Cout, Cin, 1] -> [B, Cout, n]."""
    return torch.matmul(w[:, :, 0], x) + b[None, :, None]


def linear_attention(q, k, v, eps=1e-6):
    """GATs_SuperGlue.py:69-80.  q: [B, d, H, n]; k, v: [B, d, H, m]."""
    q = F.elu(q) + 1
    k = F.elu(k) + 1
    m = v.shape[3]
    v = v / m
    KV = torch.einsum("bdhm,bqhm->bqdh", k, v)
    Z = 1 / (torch.einsum("bdhm,bdh->bhm", q, k.sum(3)) + eps)
    return torch.einsum("bdhm,bqdh,bhm->bqhm", q, KV, Z) * m


def attention_propagation(x, source, P, prefix):
    """GATs_SuperGlue.py:104-113 with MultiHeadedAttention :83-101 and MLP :116-128.
    x: [B, D, n], source: [B, D, m] -> delta [B, D, n]."""
    B, D, _ = x.shape
    d = D // NUM_HEADS
    q = _conv1x1(x, P[f"{prefix}.attn.proj.0.weight"], P[f"{prefix}.attn.proj.0.bias"])
    k = _conv1x1(source, P[f"{prefix}.attn.proj.1.weight"], P[f"{prefix}.attn.proj.1.bias"])
    v = _conv1x1(source, P[f"{prefix}.attn.proj.2.weight"], P[f"{prefix}.attn.proj.2.bias"])
    q, k, v = (t.reshape(B, d, NUM_HEADS, -1) for t in (q, k, v))        # :97 channel = d_idx*H + h
    msg = linear_attention(q, k, v).reshape(B, D, -1)                     # :100-101
    msg = _conv1x1(msg, P[f"{prefix}.attn.merge.weight"], P[f"{prefix}.attn.merge.bias"])
    y = torch.cat([x, msg], dim=1)                                        # :113
    y = _conv1x1(y, P[f"{prefix}.mlp.0.weight"], P[f"{prefix}.mlp.0.bias"])
    # nn.InstanceNorm1d(512): no affine, biased variance over the point dim, eps 1e-5 (:126)
    mean = y.mean(dim=2, keepdim=True)
    var = y.var(dim=2, unbiased=False, keepdim=True)
    y = (y - mean) / torch.sqrt(var + 1e-5)
    y = F.relu(y)
    return _conv1x1(y, P[f"{prefix}.mlp.3.weight"], P[f"{prefix}.mlp.3.bias"])


def gnn(q2d, db3d, leaves, P, hparams):
    """AttentionalGNN.forward -- GATs_SuperGlue.py:48-66.  Channel-first inputs."""
    for i, name in enumerate(GNN_LAYERS):
        p = f"gnn.layers.{i}"
        if name == "GATs":
            out = gats_layer(leaves.transpose(1, 2), db3d.transpose(1, 2), P[f"{p}.W"], P[f"{p}.a"],
                             include_self=hparams["include_self"], additional=hparams["additional"],
                             with_linear_transform=hparams["with_linear_transform"])
            db3d = out.transpose(1, 2)                                    # :51-54
        elif name == "cross":
            d0 = attention_propagation(q2d, db3d, P, p)                   # :57-58
            d1 = attention_propagation(db3d, q2d, P, p)
            q2d, db3d = q2d + d0, db3d + d1                               # :59
        else:
            d0 = attention_propagation(q2d, q2d, P, p)                    # :62-63
            d1 = attention_propagation(db3d, db3d, P, p)
            q2d, db3d = q2d + d0, db3d + d1                               # :64
    return q2d, db3d


# --------------------------------------------------------------------------
# Tail -- reference GATs_SuperGlue.py:209-237
# --------------------------------------------------------------------------
def match_tail(q2d, db3d, P, hparams):
    """final_proj + L2-normalise + dual softmax + mutual-NN.  Returns batched
    (matches0 [B,N] i64, matches1 [B,M] i64, mscores0, mscores1, conf [B,N,M],
    raw_idx0 [B,N], raw_idx1 [B,M])."""
    w, b = P["final_proj.weight"], P["final_proj.bias"]
    mq = F.normalize(_conv1x1(q2d, w, b), p=2, dim=1)                     # :209-213 (eps 1e-12)
    md = F.normalize(_conv1x1(db3d, w, b), p=2, dim=1)
    scores = torch.einsum("bdn,bdm->bnm", mq, md) / hparams["scale_factor"]   # :217
    conf = torch.softmax(scores, 1) * torch.softmax(scores, 2)            # :218
    max0, max1 = conf.max(2), conf.max(1)                                 # :220
    i0, i1 = max0.indices, max1.indices
    ar0 = torch.arange(i0.shape[1])[None]
    ar1 = torch.arange(i1.shape[1])[None]
    mutual0 = ar0 == i1.gather(1, i0)                                     # :222
    mutual1 = ar1 == i0.gather(1, i1)                                     # :223
    zero = conf.new_tensor(0)
    ms0 = torch.where(mutual0, max0.values, zero)                         # :225
    ms1 = torch.where(mutual1, ms0.gather(1, i1), zero)                   # :226
    valid0 = mutual0 & (ms0 > hparams["match_threshold"])                 # :227 (strict >)
    valid1 = mutual1 & valid0.gather(1, i1)                               # :228
    m0 = torch.where(valid0, i0, i0.new_tensor(-1))                       # :229
    m1 = torch.where(valid1, i1, i1.new_tensor(-1))                       # :230
    return m0, m1, ms0, ms1, conf, i0, i1


def forward(P, data, hparams, dtype=torch.float32):
    """Whole path, GATsSuperGlue.forward -- GATs_SuperGlue.py:179-241.

    ``data`` values: torch tensors or numpy arrays shaped like the reference's
    input dict.  Returns a dict with the batched tail outputs (the reference's
    ``pred`` is element 0 of each, :232-237) plus ``conf_matrix``."""
    def t(x):
        return torch.as_tensor(x).to(dtype)
    if hparams["match_type"] != "softmax":
        raise NotImplementedError                                         # :238-239
    k2, k3 = t(data["keypoints2d"]), t(data["keypoints3d"])
    if k2.shape[1] == 0 or k3.shape[1] == 0:                              # :195-203
        return {"matches0": torch.full(k2.shape[:-1], -1, dtype=torch.int32),
                "matches1": torch.full(k3.shape[:-1], -1, dtype=torch.int32),
                "matching_scores0": k2.new_zeros(k2.shape[:-1]),
                "matching_scores1": k3.new_zeros(k3.shape[:-1]),
                "skip_train": True}
    q2d, db3d, leaves = t(data["descriptors2d_query"]), t(data["descriptors3d_db"]), t(data["descriptors2d_db"])
    with torch.no_grad():
        q2d, db3d = gnn(q2d, db3d, leaves, P, hparams)
        m0, m1, s0, s1, conf, r0, r1 = match_tail(q2d, db3d, P, hparams)
    return {"matches0": m0, "matches1": m1, "matching_scores0": s0, "matching_scores1": s1,
            "conf_matrix": conf, "raw_indices0": r0, "raw_indices1": r1}


# --------------------------------------------------------------------------
# Offline producer: segmented mean -- reference feature_process.py:297-317
# --------------------------------------------------------------------------
def mean_descriptors(descriptors, idxs):
    """descriptors: [sum(idxs), D] float64 numpy, idxs: segment lengths -> [M, D]."""
    import numpy as np
    ends = np.cumsum(idxs)
    starts = ends - idxs
    return np.stack([descriptors[s:e].mean(axis=0) for s, e in zip(starts, ends)], 0)


def mean_scores(scores, idxs):
    """scores: [sum(idxs), 1] float64 numpy -> [M, 1] (reference feature_process.py:308-317)."""
    import numpy as np
    ends = np.cumsum(idxs)
    starts = ends - idxs
    return np.stack([scores[s:e].mean(axis=0) for s, e in zip(starts, ends)], 0)


# --------------------------------------------------------------------------
# Per-object feature construction -- reference src/utils/data_utils.py:143-205
# --------------------------------------------------------------------------
def pad_features3d_random(descriptors, scores, n_target):
    """descriptors [dim, n] f32, scores [n, 1] f32 -> padded (ones / zeros) or truncated to n_target (:143-160)."""
    import numpy as np
    descriptors = np.asarray(descriptors, np.float32)
    scores = np.asarray(scores, np.float32)
    dim, n = descriptors.shape
    if n >= n_target:                                                     # :153-155
        return descriptors[:, :n_target].copy(), scores[:n_target].copy()
    pad = n_target - n                                                    # :157-158
    return (np.concatenate([descriptors, np.ones((dim, pad), np.float32)], 1),
            np.concatenate([scores, np.zeros((pad, 1), np.float32)], 0))


def build_features3d_leaves(descriptors, scores, idxs, n_target, num_leaf):
    """Leaf selection with the dustbin column and numpy's GLOBAL RNG, one permutation per 3D point (:163-205).
    Consumes np.random exactly like the reference: seed it the same way to get the same tensors."""
    import numpy as np
    descriptors = np.asarray(descriptors, np.float32)
    scores = np.asarray(scores, np.float32)
    dim = descriptors.shape[0]
    n_points = idxs.shape[0]
    d_ext = np.concatenate([descriptors, np.ones((dim, 1), np.float32)], 1)   # :175 dustbin = all ones
    s_ext = np.concatenate([scores, np.zeros((1, 1), np.float32)], 0)        # :176
    dustbin = d_ext.shape[1] - 1
    ends = np.cumsum(idxs, axis=0)
    starts = np.insert(ends[:-1], 0, 0)
    chosen = []
    for lo, hi in zip(starts, ends):                                      # :182-192
        if num_leaf > hi - lo:
            cand = np.arange(lo, hi).tolist() + [dustbin] * (num_leaf - (hi - lo))
            chosen.append(np.random.permutation(np.array(cand)))
        else:
            chosen.append(np.random.permutation(np.arange(lo, hi))[:num_leaf])
    sel = np.concatenate(chosen, 0)
    d_out, s_out = d_ext[:, sel], s_ext[sel, :]                           # :196-197
    n_pad = n_target - n_points
    if n_pad < 0:                                                         # :199-201
        return d_out[:, :num_leaf * n_target], s_out[:num_leaf * n_target]
    return (np.concatenate([d_out, np.ones((dim, n_pad * num_leaf), np.float32)], 1),   # :203-204
            np.concatenate([s_out, np.zeros((n_pad * num_leaf, 1), np.float32)], 0))
