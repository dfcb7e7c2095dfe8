"""Counts the work of tile-culling schemes for the graph NMS on the bench workload (one image of configs[1]).
Not product code: a design aid for csrc/yb_nms_graph.cu (numbers quoted in DESIGN.md)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_path as R


def workload(conf=0.5, img=640, nc=1, seed=1234, b=0, B=2):
    grids = [img // 8, img // 16, img // 32]
    heads = []
    for s, G in enumerate(grids):
        g = torch.Generator().manual_seed(seed + s)
        heads.append(torch.randn(B, G, G, 3, 5 + nc, generator=g))
    bx, sc, cl = R.candidates([h[b:b + 1] for h in heads], R.default_anchors(), img, nc, conf)
    return bx.numpy().astype(np.float32), sc.numpy(), cl.numpy()


def spread8(v):
    v = (v | (v << 4)) & 0x0f0f
    v = (v | (v << 2)) & 0x3333
    v = (v | (v << 1)) & 0x5555
    return v


def morton(cx, cy, lo, hi, bits=8):
    qs = (2 ** bits - 1) / (hi - lo)
    fx = np.clip((cx - lo) * qs, 0, 2 ** bits - 1).astype(np.uint32)
    fy = np.clip((cy - lo) * qs, 0, 2 ** bits - 1).astype(np.uint32)
    if bits == 8:
        return spread8(fx) | (spread8(fy) << 1)
    out = np.zeros_like(fx)
    for k in range(bits):
        out |= ((fx >> k) & 1) << (2 * k)
        out |= ((fy >> k) & 1) << (2 * k + 1)
    return out


def key_current(b):
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(np.float32)
    bucket = (area.view(np.uint32) >> 24) & 0x7f
    cx, cy = (b[:, 0] + b[:, 2]) * 0.5, (b[:, 1] + b[:, 3]) * 0.5
    lo, hi = min(cx.min(), cy.min()), max(cx.max(), cy.max())
    return (bucket.astype(np.uint64) << 16) | morton(cx, cy, lo, hi)


def key_wh(b, wbits=1.0, hbits=1.0):
    """(w octave / wbits, h octave / hbits, morton)"""
    w, h = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    kw = np.floor(np.log2(np.maximum(w, 1e-3)) / wbits).astype(np.int64) + 32
    kh = np.floor(np.log2(np.maximum(h, 1e-3)) / hbits).astype(np.int64) + 32
    cx, cy = (b[:, 0] + b[:, 2]) * 0.5, (b[:, 1] + b[:, 3]) * 0.5
    lo, hi = min(cx.min(), cy.min()), max(cx.max(), cy.max())
    return (kw.astype(np.uint64) << 32) | (kh.astype(np.uint64) << 16) | morton(cx, cy, lo, hi)


def key_area(b, step=1.0):
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]))
    ka = np.floor(np.log2(np.maximum(area, 1e-3)) / step).astype(np.int64) + 32
    cx, cy = (b[:, 0] + b[:, 2]) * 0.5, (b[:, 1] + b[:, 3]) * 0.5
    lo, hi = min(cx.min(), cy.min()), max(cx.max(), cy.max())
    return (ka.astype(np.uint64) << 16) | morton(cx, cy, lo, hi)


def stats(b, n_per):
    """bbox + area range (+ w/h ranges, centre ranges) per group of n_per consecutive boxes"""
    M = b.shape[0]
    G = (M + n_per - 1) // n_per
    pad = G * n_per - M
    def red(v, fn, fill):
        vv = np.concatenate([v, np.full(pad, fill, v.dtype)]).reshape(G, n_per)
        return fn(vv, axis=1)
    w, h = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    a = w * h
    cx, cy = (b[:, 0] + b[:, 2]) * 0.5, (b[:, 1] + b[:, 3]) * 0.5
    inf = np.float32(np.inf)
    return dict(x1=red(b[:, 0], np.min, inf), y1=red(b[:, 1], np.min, inf), x2=red(b[:, 2], np.max, -inf), y2=red(b[:, 3], np.max, -inf),
                amin=red(a, np.min, inf), amax=red(a, np.max, -inf), wmin=red(w, np.min, inf), wmax=red(w, np.max, -inf),
                hmin=red(h, np.min, inf), hmax=red(h, np.max, -inf), cx1=red(cx, np.min, inf), cx2=red(cx, np.max, -inf),
                cy1=red(cy, np.min, inf), cy2=red(cy, np.max, -inf))


def group_vs_group(S, T, t2, mode):
    """boolean (len S, len T): can any box of group s have IoU > thr with any box of group t"""
    ox = np.minimum(S['x2'][:, None], T['x2'][None]) - np.maximum(S['x1'][:, None], T['x1'][None])
    oy = np.minimum(S['y2'][:, None], T['y2'][None]) - np.maximum(S['y1'][:, None], T['y1'][None])
    ok = (ox > 0) & (oy > 0) & (S['amax'][:, None] >= t2 * T['amin'][None]) & (T['amax'][None] >= t2 * S['amin'][:, None])
    if mode == 'wh':
        # IoU <= min(w)/max(w) * min(h)/max(h): best case ratio per dim from the ranges
        def best_ratio(lo1, hi1, lo2, hi2):
            # max over a in [lo1,hi1], b in [lo2,hi2] of min(a,b)/max(a,b)
            overlap = (np.minimum(hi1, hi2) >= np.maximum(lo1, lo2))
            r = np.where(overlap, 1.0, np.where(hi1 < lo2, hi1 / lo2, hi2 / lo1))
            return r
        rw = best_ratio(S['wmin'][:, None], S['wmax'][:, None], T['wmin'][None], T['wmax'][None])
        rh = best_ratio(S['hmin'][:, None], S['hmax'][:, None], T['hmin'][None], T['hmax'][None])
        ok &= rw * rh >= t2
        # centre distance: overlap_x = (w1+w2)/2 - |dx| >= t2*max(w1,w2) needs |dx| <= (w1+w2)/2 - t2*max(w1,w2) <= wmaxS/2+wmaxT/2 - t2*max(wminS,wminT)... use loose bound
        dx = np.maximum(0, np.maximum(S['cx1'][:, None] - T['cx2'][None], T['cx1'][None] - S['cx2'][:, None]))
        dy = np.maximum(0, np.maximum(S['cy1'][:, None] - T['cy2'][None], T['cy1'][None] - S['cy2'][:, None]))
        wmx = np.maximum(S['wmax'][:, None], T['wmax'][None]); wmn = np.minimum(S['wmax'][:, None], T['wmax'][None])
        hmx = np.maximum(S['hmax'][:, None], T['hmax'][None]); hmn = np.minimum(S['hmax'][:, None], T['hmax'][None])
        ok &= (dx <= (wmx + wmn) / 2 - t2 * np.maximum(S['wmin'][:, None], T['wmin'][None]))
        ok &= (dy <= (hmx + hmn) / 2 - t2 * np.maximum(S['hmin'][:, None], T['hmin'][None]))
    return ok


def simulate(b, key, thr=0.4, tile=32, sub=8, mode='bbox', row_level=True, verbose=True):
    t2 = thr * (1 - 2 ** -10)
    order = np.argsort(key, kind='stable')
    sb = b[order]
    M = sb.shape[0]
    Ts, Ss = stats(sb, tile), stats(sb, sub)
    nT, nS = len(Ts['x1']), len(Ss['x1'])
    k = tile // sub
    # level 1: tile pairs J >= I
    l1 = group_vs_group(Ts, Ts, t2, mode)
    l1 &= np.triu(np.ones((nT, nT), bool))
    # level 2: tile I vs subtiles of surviving J
    l2 = group_vs_group(Ts, Ss, t2, mode) & np.repeat(l1, k, axis=1)[:, :nS]
    # level 3: rows vs sub-tiles
    R1 = stats(sb, 1)
    items = 0
    tests = 0
    chunks = 0
    for I in range(nT):
        cols = np.nonzero(l2[I])[0]
        if len(cols) == 0:
            continue
        chunks += (len(cols) + 3) // 4
        rows = slice(I * tile, min(M, (I + 1) * tile))
        Rr = {kk: v[rows] for kk, v in R1.items()}
        Sc = {kk: v[cols] for kk, v in Ss.items()}
        if row_level:
            ox = np.minimum(Rr['x2'][:, None], Sc['x2'][None]) - np.maximum(Rr['x1'][:, None], Sc['x1'][None])
            oy = np.minimum(Rr['y2'][:, None], Sc['y2'][None]) - np.maximum(Rr['y1'][:, None], Sc['y1'][None])
            rw, rh, rS = Rr['wmax'][:, None], Rr['hmax'][:, None], Rr['amax'][:, None]
            ok = (ox > 0) & (oy > 0) & (ox >= t2 * rw) & (oy >= t2 * rh) & (Sc['amax'][None] >= t2 * rS) & (rS >= t2 * Sc['amin'][None])
            if mode == 'wh':
                ok &= group_vs_group(Rr, Sc, t2, 'wh')
            items += int(ok.sum())
        else:
            items += (rows.stop - rows.start) * len(cols)
    tests = items * sub
    # true edges
    res = dict(M=M, tile_pairs=int(l1.sum()), sub_pairs=int(l2.sum()), chunks=chunks, items=items, tests=tests)
    if verbose:
        print(res, "tests/box=%.1f" % (tests / M))
    return res


def true_edges(b, thr=0.4):
    M = b.shape[0]
    n = 0
    a = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    for i0 in range(0, M, 1024):
        x = b[i0:i0 + 1024]
        iw = np.clip(np.minimum(x[:, None, 2], b[None, :, 2]) - np.maximum(x[:, None, 0], b[None, :, 0]), 0, None)
        ih = np.clip(np.minimum(x[:, None, 3], b[None, :, 3]) - np.maximum(x[:, None, 1], b[None, :, 1]), 0, None)
        inter = iw * ih
        iou = inter / (a[i0:i0 + 1024, None] + a[None] - inter)
        n += int((iou > thr).sum()) - x.shape[0]
    return n // 2


if __name__ == "__main__":
    conf = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
    b, sc, cl = workload(conf)
    print("M", b.shape[0], "edges", true_edges(b))
    print("current key, bbox stats:")
    simulate(b, key_current(b))
    print("current key, no row culling (tile-8 x row):")
    simulate(b, key_current(b), row_level=False)
    for step in (2.0, 1.0, 0.5):
        print("area step", step)
        simulate(b, key_area(b, step))
    for wb in (2.0, 1.0):
        print("wh key", wb, "bbox"); simulate(b, key_wh(b, wb, wb))
        print("wh key", wb, "wh stats"); simulate(b, key_wh(b, wb, wb), mode='wh')
    print("current key + wh stats"); simulate(b, key_current(b), mode='wh')


def per_tile_cost(b, key, thr=0.4, tile=32, sub=8):
    t2 = thr * (1 - 2 ** -10)
    order = np.argsort(key, kind='stable')
    sb = b[order]
    Ts, Ss = stats(sb, tile), stats(sb, sub)
    nT, nS = len(Ts['x1']), len(Ss['x1'])
    k = tile // sub
    l1 = group_vs_group(Ts, Ts, t2, 'bbox') & np.triu(np.ones((nT, nT), bool))
    l2 = group_vs_group(Ts, Ss, t2, 'bbox') & np.repeat(l1, k, axis=1)[:, :nS]
    return l1.sum(1), l2.sum(1)


def hilbert(cx, cy, lo, hi, bits=8):
    n = 1 << bits
    qs = (n - 1) / (hi - lo)
    x = np.clip((cx - lo) * qs, 0, n - 1).astype(np.int64)
    y = np.clip((cy - lo) * qs, 0, n - 1).astype(np.int64)
    d = np.zeros_like(x)
    s = n // 2
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate
        m = ry == 0
        flip = m & (rx == 1)
        x = np.where(flip, s - 1 - (x & (s - 1)) + (x & ~(s - 1)) * 0, x)  # placeholder, fixed below
        s //= 2
    return d


def hilbert_d(cx, cy, lo, hi, bits=8):
    """classic xy2d"""
    n = 1 << bits
    qs = (n - 1) / (hi - lo)
    x = np.clip((cx - lo) * qs, 0, n - 1).astype(np.int64)
    y = np.clip((cy - lo) * qs, 0, n - 1).astype(np.int64)
    d = np.zeros_like(x)
    s = n // 2
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        m = ry == 0
        fl = m & (rx == 1)
        x2 = np.where(fl, n - 1 - x, x)
        y2 = np.where(fl, n - 1 - y, y)
        x, y = np.where(m, y2, x2), np.where(m, x2, y2)
        s //= 2
    return d.astype(np.uint64)


def key_variant(b, curve='morton', alternate=False):
    area = ((b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(np.float32)
    bucket = ((area.view(np.uint32) >> 24) & 0x7f).astype(np.uint64)
    cx, cy = (b[:, 0] + b[:, 2]) * 0.5, (b[:, 1] + b[:, 3]) * 0.5
    lo, hi = min(cx.min(), cy.min()), max(cx.max(), cy.max())
    c = morton(cx, cy, lo, hi).astype(np.uint64) if curve == 'morton' else hilbert_d(cx, cy, lo, hi)
    if alternate:
        c = np.where(bucket & 1, np.uint64(0xffff) - c, c)
    return (bucket << np.uint64(16)) | c
