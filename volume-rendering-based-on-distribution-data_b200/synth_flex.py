"""Synthetic, self-consistent span store for the flexible-block chain (test / benchmark data).

The reference's span tables come from data files that do not ship.  This generator builds them from a seeded
raw volume the way an encoder would: every power-of-two-aligned box ("span") gets the true normalised 64-bin
histogram of the raw voxels it covers; boxes of >= 8 voxels are stored as LOSSLESS fractal codes (a template,
a shift, a flip, and one sparse error per bin that differs), smaller boxes as sparse "simple" histograms with
0-based coordinates (volumeRender_kernel.cu:1464-1471).  Because the codes are lossless, the chain's output
can be checked against histograms counted directly from the raw volume."""
import numpy as np

BINS = 64


def raw_volume(seed, R):
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(R), np.arange(R), np.arange(R), indexing="ij")
    smooth = 110 + 70 * np.sin(x * 2.1 * np.pi / R) * np.cos(y * 1.3 * np.pi / R) + 50 * np.sin(z * 1.7 * np.pi / R + 0.5)
    return np.clip(smooth + rng.normal(0, 12, (R, R, R)), 0, 255).astype(np.uint8)


def _integral(raw):
    R = raw.shape[0]
    onehot = np.zeros((R, R, R, BINS), np.int32)
    np.put_along_axis(onehot, (raw.astype(np.int64) >> 2)[..., None], 1, axis=3)
    ih = np.zeros((R + 1, R + 1, R + 1, BINS), np.int64)
    ih[1:, 1:, 1:] = onehot.cumsum(0).cumsum(1).cumsum(2)
    return ih


def box_counts(ih, lo, hi):
    """counts of the 1-based inclusive box lo..hi (x, y, z order)"""
    (lx, ly, lz), (hx, hy, hz) = lo, hi
    a, b, c = lx - 1, ly - 1, lz - 1
    return (ih[hz, hy, hx] - ih[c, hy, hx] - ih[hz, b, hx] - ih[hz, hy, a] + ih[c, b, hx] + ih[c, hy, a] + ih[hz, b, a] - ih[c, b, a])


def dyadic_intervals(R, needed=None):
    out = []
    size = 1
    while size <= R:
        for k in range(R // size):
            out.append((k * size + 1, (k + 1) * size))
        size *= 2
    if needed is not None:
        keep = set()
        for x in needed:
            i = 0
            while x:
                if x & (1 << i):
                    keep.add((x - (1 << i) + 1, x)); x &= ~(1 << i)
                i += 1
        out = [iv for iv in out if iv in keep]
    return out


def make_tables(seed, R, n_templates=12, block=None):
    """Span store of an R^3 raw volume.  With `block`, only the spans a query with that block size needs."""
    rng = np.random.default_rng(seed + 1)
    raw = raw_volume(seed, R)
    ih = _integral(raw)
    needed = None
    if block:
        needed = set()
        nb = (R + block - 1) // block
        for i in range(nb):
            needed.add(1 + i * block); needed.add(R if i == nb - 1 else (i + 1) * block)
    iv = dyadic_intervals(R, needed)
    templates = rng.random((n_templates, BINS)).astype(np.float32) ** 4
    templates /= templates.sum(1, keepdims=True)
    templates = templates.astype(np.float32)
    f_low, f_high, f_code, f_err, s_low, s_high, s_cnt, s_hist = [], [], [], [], [], [], [], []
    for (zl, zh) in iv:
        for (yl, yh) in iv:
            for (xl, xh) in iv:
                size = (xh - xl + 1) * (yh - yl + 1) * (zh - zl + 1)
                cnt = box_counts(ih, (xl, yl, zl), (xh, yh, zh))
                h = (cnt / size).astype(np.float32)
                if size >= 8:
                    hh = int(rng.integers(0, 1 << 30))
                    tid, shift, flip = hh % n_templates, (hh >> 8) % (BINS + 1), (hh >> 20) & 1
                    src = templates[tid][::-1] if flip else templates[tid]
                    cur = np.roll(src, shift % BINS)
                    diff = np.nonzero(h != cur)[0]
                    err = np.zeros((BINS, 2), np.float32)
                    err[:len(diff), 0] = diff
                    err[:len(diff), 1] = (h[diff].astype(np.float64) - cur[diff].astype(np.float64)).astype(np.float32)
                    f_low.append((xl, yl, zl, 0)); f_high.append((xh, yh, zh, 0))
                    f_code.append((tid, shift, flip, len(diff))); f_err.append(err)
                else:
                    nz = np.nonzero(cnt)[0]
                    e = np.zeros((BINS, 2), np.float32)
                    e[:len(nz), 0] = nz; e[:len(nz), 1] = h[nz]
                    s_low.append((xl - 1, yl - 1, zl - 1, 0)); s_high.append((xh - 1, yh - 1, zh - 1, 0))
                    s_cnt.append(len(nz)); s_hist.append(e)
    i32 = lambda a, w: np.array(a, np.int32).reshape(-1, w)
    return {"raw_dims": (R, R, R), "bins": BINS, "raw": raw, "integral": ih,
            "span_low": i32(f_low, 4), "span_high": i32(f_high, 4), "codebook": i32(f_code, 4),
            "errors": np.array(f_err, np.float32).reshape(-1, BINS, 2),
            "simple_low": i32(s_low, 4), "simple_high": i32(s_high, 4), "simple_count": np.array(s_cnt, np.int32),
            "simple_hist": np.array(s_hist, np.float32).reshape(-1, BINS, 2), "templates": templates}


def expected_blocks(tables, block):
    """Block statistics counted directly from the raw volume, following what the reference's chain really
    computes: (a) the lower corners use `low`, not `low - 1` (volumeRender_kernel.cu:1157-1227), so in x and y a
    block covers (low, high]; (b) the corner signs are +0 +3 +4 +7 -1 -2 -5 -6 (:1043-1046), i.e. the 2-D
    inclusion-exclusion at z = low and at z = high ADDED, not subtracted: the block's histogram is the xy-box
    over z in [1, high] plus the same box over z in [1, low]."""
    R = tables["raw_dims"][0]
    ih = tables["integral"]
    nb = (R + block - 1) // block
    out = np.zeros((nb * nb * nb, 4))
    bw = 255.0 / BINS
    c = bw * np.arange(BINS) + bw / 2
    for b in range(nb ** 3):
        q = (b % nb, (b // nb) % nb, b // (nb * nb))
        lo = [1 + k * block for k in q]; hi = [R if k == nb - 1 else (k + 1) * block for k in q]
        if lo[0] + 1 > hi[0] or lo[1] + 1 > hi[1]:
            continue                                          # degenerate (empty) block: total 0, stats 0
        cnt = (box_counts(ih, (lo[0] + 1, lo[1] + 1, 1), (hi[0], hi[1], hi[2])) +
               box_counts(ih, (lo[0] + 1, lo[1] + 1, 1), (hi[0], hi[1], lo[2]))).astype(np.float64)
        p = cnt / cnt.sum()
        m = (p * c).sum()
        out[b, 0] = m; out[b, 1] = (p * (c - m) ** 2).sum()
        nzp = p[p > 0]
        out[b, 2] = -(nzp * np.log2(nzp)).sum() / 6.0
    return out
