"""numpy emulation of the extractor's DEVICE algorithms (csrc/superpoint.cu), used by the CPU tests to check the design
against torch before a GPU is involved: zero-bordered pixel-grid layout + 9 row-shifted GEMMs for a 3x3 convolution, the
weight packing order, 2x2 pooling on the grid, the tiled NMS with its shrinking valid margins, ordered compaction + rank-by-
counting top-k, and the bilinear descriptor sampling.  Mirrors the kernels' index arithmetic line by line (fp32 numpy instead
of the fp16-split tensor-core product)."""
import numpy as np


def round_up(x, m):
    return (x + m - 1) // m * m


def stage(H, W, s):
    h, w = H >> s, W >> s
    return h, w, round_up((h + 2) * (w + 2), 256)


def to_grid(x):
    """[B, C, h, w] -> rows [B * P, C] of the zero-bordered grid."""
    B, C, h, w = x.shape
    _, _, P = stage(h, w, 0)
    g = np.zeros((B, P, C), np.float32)
    pad = np.zeros((B, h + 2, w + 2, C), np.float32)
    pad[:, 1:-1, 1:-1] = x.transpose(0, 2, 3, 1)
    g[:, :(h + 2) * (w + 2)] = pad.reshape(B, -1, C)
    return g.reshape(B * P, C)


def from_grid(rows, B, C, h, w):
    _, _, P = stage(h, w, 0)
    g = rows.reshape(B, P, C)[:, :(h + 2) * (w + 2)].reshape(B, h + 2, w + 2, C)
    return g[:, 1:-1, 1:-1].transpose(0, 3, 1, 2)


def pack_conv(weight):
    """sp_pack: [cout, cin, k, k] -> B operand [cout, taps * cin], reduction index tap * cin + c, tap = ky * k + kx."""
    cout, cin, k, _ = weight.shape
    return weight.transpose(0, 2, 3, 1).reshape(cout, k * k * cin)


def conv_grid(rows, wpacked, bias, h, w, P, cin, relu=True):
    """gemm_tc.cu EPI_CONV with GemmProblem::taps: out[p] = sum_t A[p + off_t] . W_t^T, rows outside [0, R) read as zero (TMA
    fill), border / padding rows of the output written as zero."""
    R = rows.shape[0]
    taps = wpacked.shape[1] // cin
    offs = [(ky - 1) * (w + 2) + (kx - 1) for ky in range(3) for kx in range(3)] if taps == 9 else [0]
    acc = np.zeros((R, wpacked.shape[0]), np.float64)
    for t, off in enumerate(offs):
        src = np.zeros_like(rows)
        lo, hi = max(0, -off), min(R, R - off)
        src[lo:hi] = rows[lo + off:hi + off]
        acc += src.astype(np.float64) @ wpacked[:, t * cin:(t + 1) * cin].T.astype(np.float64)
    out = acc + bias[None]
    if relu:
        out = np.maximum(out, 0)
    q = np.arange(R) % P
    yy, xx = q // (w + 2), q % (w + 2)
    ok = (yy >= 1) & (yy <= h) & (xx >= 1) & (xx <= w)
    out[~ok] = 0
    return out.astype(np.float32)


def pool_grid(rows, B, C, hin, win):
    """sp_pool2x2."""
    _, _, Pin = stage(hin, win, 0)
    ho, wo = hin // 2, win // 2
    _, _, Pout = stage(ho, wo, 0)
    out = np.zeros((B * Pout, C), np.float32)
    r = np.arange(B * Pout)
    b, q = r // Pout, r % Pout
    y, x = q // (wo + 2) - 1, q % (wo + 2) - 1
    ok = (y >= 0) & (y < ho) & (x >= 0) & (x < wo)
    best = None
    for d in range(4):
        src = b[ok] * Pin + (2 * y[ok] + (d >> 1) + 1) * (win + 2) + (2 * x[ok] + (d & 1) + 1)
        v = rows[src]
        best = v if best is None else np.maximum(best, v)
    out[ok] = best
    return out


def _pool_sep(inp, D, margin, r):
    out = np.full((D, D), np.nan, np.float32)
    lo, hi = margin, D - margin
    tmp = np.full((D, D), np.nan, np.float32)
    for x in range(lo + r, hi - r):
        tmp[lo:hi, x] = inp[lo:hi, x - r:x + r + 1].max(1)
    for y in range(lo + r, hi - r):
        out[y, lo + r:hi - r] = tmp[y - r:y + r + 1, lo + r:hi - r].max(0)
    return out


def nms_tiled(scores, r, tile=32):
    """sp_nms: one tile at a time with a 5 r halo and shrinking valid margins."""
    H, W = scores.shape
    out = np.zeros_like(scores)
    halo = 5 * r
    D = tile + 2 * halo
    for ty in range(0, H, tile):
        for tx in range(0, W, tile):
            S = np.full((D, D), -np.inf, np.float32)
            ys, xs = np.arange(ty - halo, ty - halo + D), np.arange(tx - halo, tx - halo + D)
            vy, vx = (ys >= 0) & (ys < H), (xs >= 0) & (xs < W)
            S[np.ix_(vy, vx)] = scores[np.ix_(ys[vy], xs[vx])]
            inside = S != -np.inf
            P = _pool_sep(S, D, 0, r)
            margin = r
            Mk = np.zeros((D, D), np.float32)
            sl = slice(margin, D - margin)
            Mk[sl, sl] = ((S[sl, sl] == P[sl, sl]) & inside[sl, sl]).astype(np.float32)
            for _ in range(2):
                Sup = _pool_sep(Mk, D, margin, r)
                margin += r
                sl = slice(margin, D - margin)
                SS = np.full((D, D), np.nan, np.float32)
                SS[sl, sl] = np.where(inside[sl, sl], np.where(Sup[sl, sl] > 0, 0.0, S[sl, sl]), -np.inf)
                P = _pool_sep(SS, D, margin, r)
                margin += r
                sl = slice(margin, D - margin)
                new = (~(Sup[sl, sl] > 0)) & (SS[sl, sl] == P[sl, sl]) & inside[sl, sl]
                Mk[sl, sl] = np.where(new, 1.0, Mk[sl, sl])
            assert margin == halo
            t = S[halo:halo + tile, halo:halo + tile]
            m = Mk[halo:halo + tile, halo:halo + tile] == 1
            hh, ww = min(tile, H - ty), min(tile, W - tx)
            out[ty:ty + hh, tx:tx + ww] = np.where(m, t, 0)[:hh, :ww]
    return out


def select(nms, thr, border, k):
    """sp_count / sp_compact / sp_emit: ordered candidates, rank-by-counting top-k (equal scores: lower pixel index first)."""
    H, W = nms.shape
    pix = np.arange(H * W)
    y, x = pix // W, pix % W
    s = nms.reshape(-1)
    keep = (s > thr) & (y >= border) & (y < H - border) & (x >= border) & (x < W - border)
    ci, cs = pix[keep], s[keep]
    total = len(ci)
    if k >= 0 and total > k:
        rank = np.array([int(((cs > cs[i]) | ((cs == cs[i]) & (ci < ci[i]))).sum()) for i in range(total)])
        sel = rank < k
        order = np.empty(k, np.int64)
        order[rank[sel]] = np.nonzero(sel)[0]
        ci, cs = ci[order], cs[order]
    kp = np.stack([ci % W, ci // W], 1).astype(np.float32)
    return kp, cs


def sample(ddesc_rows, h, w, P, kpts, align_corners):
    """sp_sample for image 0: ddesc rows [P, 256] of the coarse grid (un-normalised convDb output)."""
    f = np.float32
    s = f(8)
    out = np.zeros((256, len(kpts)), np.float32)
    for i, (kx, ky) in enumerate(kpts.astype(np.float32)):
        gx = (kx - s / f(2) + f(0.5)) / f(w * 8 - 4 - 0.5)
        gy = (ky - s / f(2) + f(0.5)) / f(h * 8 - 4 - 0.5)
        gx, gy = gx * f(2) - f(1), gy * f(2) - f(1)
        if align_corners:
            ix, iy = ((gx + f(1)) / f(2)) * f(w - 1), ((gy + f(1)) / f(2)) * f(h - 1)
        else:
            ix, iy = ((gx + f(1)) * f(w) - f(1)) / f(2), ((gy + f(1)) * f(h) - f(1)) / f(2)
        x0, y0 = int(np.floor(ix)), int(np.floor(iy))
        wx1, wy1 = ix - f(x0), iy - f(y0)
        wgt = [(f(1) - wx1) * (f(1) - wy1), wx1 * (f(1) - wy1), (f(1) - wx1) * wy1, wx1 * wy1]
        acc = np.zeros(256, np.float32)
        for d in range(4):
            yy, xx = y0 + (d >> 1), x0 + (d & 1)
            if yy < 0 or yy >= h or xx < 0 or xx >= w:
                continue
            v = ddesc_rows[(yy + 1) * (w + 2) + xx + 1]
            acc += v / max(np.sqrt((v * v).sum()), 1e-12) * wgt[d]
        out[:, i] = acc / max(np.sqrt((acc * acc).sum()), 1e-12)
    return out
