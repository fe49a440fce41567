"""CPU oracle of the SuperPoint extractor -- TEST INFRASTRUCTURE, never imported by the product path.

Functional torch-CPU restatement of ``SuperPoint.forward``
(reference ``src/models/extractors/SuperPoint/superpoint.py``), each step citing the lines it follows.  Pinned by
``tests/test_oracle_golden.py`` to fixtures produced by the unmodified reference module
(``tests/golden/make_golden.py``: ``superpoint_*.npz``).  Only ``tests/``, ``tools/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU legs may import this file.

``align_corners``: the reference passes ``align_corners=True`` to ``grid_sample`` when ``int(torch.__version__[2]) > 2``
(superpoint.py:86) -- true for its pinned torch 1.8.0 ('1.8.0'[2] == '8'), false for this image's torch 2.11 ('2.11'[2] == '1').
Both branches are restated and pinned; the product defaults to the pinned environment's behaviour (True).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

DEFAULT_CONFIG = {
    # superpoint.py:97-103
    "descriptor_dim": 256,
    "nms_radius": 4,
    "keypoint_threshold": 0.005,
    "max_keypoints": -1,
    "remove_borders": 4,
}

ENCODER = ["conv1a", "conv1b", "pool", "conv2a", "conv2b", "pool", "conv3a", "conv3b", "pool", "conv4a", "conv4b"]


def params_from_numpy(sd):
    return {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in sd.items()}


def _conv(p, name, x, relu=True):
    w = p[name + ".weight"]
    y = F.conv2d(x, w, p[name + ".bias"], stride=1, padding=w.shape[-1] // 2)
    return F.relu(y) if relu else y


def encoder(p, inp, upto=None):
    """superpoint.py:142-152.  ``upto`` = index into ENCODER to stop after (intermediate activations for the layer tests)."""
    x = inp
    for i, name in enumerate(ENCODER):
        x = F.max_pool2d(x, kernel_size=2, stride=2) if name == "pool" else _conv(p, name, x)
        if upto is not None and i == upto:
            break
    return x


def dense_scores(p, x):
    """superpoint.py:155-160: score head, softmax over 65 channels, dustbin dropped, 8x8 pixel shuffle."""
    logits = _conv(p, "convPb", _conv(p, "convPa", x), relu=False)
    s = F.softmax(logits, 1)[:, :-1]
    b, _, h, w = s.shape
    s = s.permute(0, 2, 3, 1).reshape(b, h, w, 8, 8)
    return s.permute(0, 1, 3, 2, 4).reshape(b, h * 8, w * 8)


def simple_nms(scores, r):
    """superpoint.py:47-61."""
    def mp(x):
        return F.max_pool2d(x, kernel_size=2 * r + 1, stride=1, padding=r)
    zeros = torch.zeros_like(scores)
    mask = scores == mp(scores)
    for _ in range(2):
        supp = mp(mask.float()) > 0
        ss = torch.where(supp, zeros, scores)
        new = ss == mp(ss)
        mask = mask | (new & ~supp)
    return torch.where(mask, scores, zeros)


def select_keypoints(nms, cfg):
    """superpoint.py:163-177 for one image: threshold, border removal, top-k, (h, w) -> (x, y)."""
    H, W = nms.shape
    k = torch.nonzero(nms > cfg["keypoint_threshold"])
    s = nms[tuple(k.t())]
    b = cfg["remove_borders"]
    m = (k[:, 0] >= b) & (k[:, 0] < H - b) & (k[:, 1] >= b) & (k[:, 1] < W - b)      # :64-69
    k, s = k[m], s[m]
    if cfg["max_keypoints"] >= 0 and cfg["max_keypoints"] < len(k):                  # :72-76
        s, idx = torch.topk(s, cfg["max_keypoints"], dim=0)
        k = k[idx]
    return torch.flip(k, [1]).float(), s


def dense_descriptors(p, x):
    """superpoint.py:180-181."""
    d = _conv(p, "convDb", _conv(p, "convDa", x), relu=False)
    return F.normalize(d, p=2, dim=1)


def sample_descriptors(kpts, desc, s=8, align_corners=True):
    """superpoint.py:79-92 for one image: kpts [n, 2] (x, y), desc [C, h, w] -> [C, n]."""
    c, h, w = desc.shape
    k = kpts - s / 2 + 0.5
    k = k / torch.tensor([(w * s - s / 2 - 0.5), (h * s - s / 2 - 0.5)]).to(k)[None]
    k = k * 2 - 1
    d = F.grid_sample(desc[None], k.view(1, 1, -1, 2), mode="bilinear", align_corners=align_corners)
    return F.normalize(d.reshape(1, c, -1), p=2, dim=1)[0]


def forward(p, image, config=None, align_corners=True):
    """SuperPoint.forward (superpoint.py:140-197).  image fp32 [B, 1, H, W] (numpy or torch)."""
    cfg = {**DEFAULT_CONFIG, **(config or {})}
    if cfg["max_keypoints"] == 0 or cfg["max_keypoints"] < -1:
        raise ValueError('"max_keypoints" must be positive or "-1"')      # :133-135
    inp = torch.as_tensor(image).float()
    with torch.no_grad():
        x = encoder(p, inp)
        scores = dense_scores(p, x)
        nms = simple_nms(scores, cfg["nms_radius"])
        dd = dense_descriptors(p, x)
        kp, sc, de = [], [], []
        for b in range(inp.shape[0]):
            k, s = select_keypoints(nms[b], cfg)
            kp.append(k)
            sc.append(s)
            de.append(sample_descriptors(k, dd[b], 8, align_corners))
    return {"keypoints": kp, "scores": sc, "descriptors": de, "dense_scores": scores, "nms": nms}


def compare_keypoints(mine_k, mine_s, ref_k, ref_s, tol, max_keypoints=-1):
    """Key-point parity under a score tolerance (checker for tests/ and bench.py).

    Row-major output (no top-k): the key points must be identical, in order.  Top-k output (len(ref) == max_keypoints): the
    order among candidates whose scores differ by less than `tol` -- and membership for candidates within `tol` of the k-th
    score -- is decided by fp32 noise in the reference itself, so the check is:
      * the descending score sequences agree element-wise within tol;
      * every key point present on both sides carries the same score within tol;
      * a key point present on one side only has a score within tol of the cut (the k-th score).
    Returns a dict of diagnostics; raises AssertionError with a precise message otherwise."""
    mine_k, mine_s, ref_k, ref_s = (np.asarray(a) for a in (mine_k, mine_s, ref_k, ref_s))
    topk = max_keypoints >= 0 and len(ref_k) == max_keypoints
    if not topk:
        assert mine_k.shape == ref_k.shape, f"{len(mine_k)} key points, reference {len(ref_k)}"
        assert np.array_equal(mine_k, ref_k), "key points differ (row-major selection must be bit-identical)"
        err = float(np.abs(mine_s - ref_s).max()) if len(ref_s) else 0.0
        assert err <= tol, f"score error {err:.2e}"
        return {"mode": "row-major", "identical": True, "max_score_err": err, "n": int(len(ref_k))}
    assert len(mine_k) == len(ref_k), f"{len(mine_k)} key points, reference {len(ref_k)}"
    seq_err = float(np.abs(mine_s - ref_s).max())
    assert seq_err <= tol, f"descending score sequences differ by {seq_err:.2e}"
    key = lambda k: (int(k[0]), int(k[1]))      # noqa: E731
    rs = {key(k): float(s) for k, s in zip(ref_k, ref_s)}
    ms = {key(k): float(s) for k, s in zip(mine_k, mine_s)}
    cut = float(ref_s[-1])
    both = [abs(ms[k] - rs[k]) for k in ms if k in rs]
    assert max(both) <= tol, f"score of a common key point differs by {max(both):.2e}"
    only = [ms[k] for k in ms if k not in rs] + [rs[k] for k in rs if k not in ms]
    assert all(abs(s - cut) <= tol for s in only), f"key points on one side only are not at the cut: {only[:4]} vs cut {cut}"
    same_pos = int((mine_k == ref_k).all(1).sum())
    return {"mode": "top-k", "identical": bool(same_pos == len(ref_k)), "same_position": same_pos, "n": int(len(ref_k)),
            "one_sided": len(only) // 2, "max_score_err": float(max(both)), "seq_err": seq_err}
