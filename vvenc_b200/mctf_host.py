"""Host-side replay of the MCTF motion search control, on top of batched error tables.

MCTF::estimateLumaLn (CommonLib/MCTF.cpp:1166-1327) decides with `error < best.error` chains over a few candidate sets; the errors themselves are what
the GPU library returns in bulk (vvb_mctf_search_grid: whole search grids per block, vvb_mctf_error_batch: explicit candidates, vvb_mctf_calc_var).
`estimate_level` reproduces one pyramid level (MCTF::motionEstimationLuma, :1329-1397) for the three search patterns (MCTFSpeed 0 / 1-2 / 3-4, :598-599):

  stage A  predictors: the 3x3 neighbourhood of the coarser level's field (scaled by `factor`) and the zero vector            (:1191-1214)
  stage B  integer grid around trunc(best / 16): range 8 without a coarser level, 5 with one, none when doubleRes              (:1216-1228)
  stage C  doubleRes only: 7x7 grid step 4, 3x3 step 2, 3x3 step 1 around the running best, centre skipped                    (:1229-1287)
  stage D  the final vectors of the block above and of the block to the left                                                  (:1288-1306)
  stage E  doubleRes only: error -> 20 * ((err * s + 5) / (var * s + 5)) + mse / 50, rmsme, overlap                           (:1308-1321)

Stages A-C do not look at neighbouring blocks, so they are evaluated for all blocks at once (one provider call per stage); stage D follows the
dependency on the upper and left neighbours along anti-diagonals (all blocks with the same bx + by in one call).  Early exits of the reference's error
functions never change a decision (a partial sum is only returned when it already exceeds the best), so full sums give identical fields.

A *provider* supplies the numbers:
    grid(blocks, step, radius)  -> int32 [n][2r+1][2r+1]   blocks: structured array x, y, mvx, mvy (centre, 1/16 pel), w, h   (vvb_mctf_search_grid)
    errors(cands)               -> int32 [n]               cands : structured array x, y, mvx, mvy, w, h                        (vvb_mctf_error_batch)
    calc_var(blocks)            -> float64 [n]                                                                                 (vvb_mctf_calc_var)
`EngineProvider` wraps a vvenc_b200.CostEngine; the tests drive the same replay with a provider backed by the CPU oracle and compare the field with
the reference's own motionEstimationLuma.
"""
import numpy as np

CAND_DT = np.dtype([('x', '<i4'), ('y', '<i4'), ('mvx', '<i4'), ('mvy', '<i4'), ('w', '<u2'), ('h', '<u2')])
INT_MAX = 2 ** 31 - 1              # MotionVector() starts with error = INT_LEAST32_MAX (MCTF.h:79)


class EngineProvider:
    """numbers from the GPU library: org_plane / ref_plane are resident plane ids of a CostEngine"""

    def __init__(self, engine, org_plane, ref_plane, low_res_filter=False):
        self.eng = engine; self.org = org_plane; self.ref = ref_plane; self.low = bool(low_res_filter)

    def grid(self, blocks, step, radius):
        return self.eng.mctf_search_grid(self.org, self.ref, blocks, step, radius, self.low)

    def errors(self, cands):
        return self.eng.mctf_error_batch(self.org, self.ref, cands, self.low)

    def calc_var(self, blocks):
        return self.eng.mctf_calc_var(self.org, blocks)


def _trunc_div16(v):
    """C integer division by 16 (truncation toward zero), as `prevBest.y / m_motionVectorFactor`"""
    v = np.asarray(v, dtype=np.int64)
    return np.where(v >= 0, v // 16, -((-v) // 16)).astype(np.int32)


def _offset_table(provider, cands, cx, cy, offs):
    """errors of the vectors (cx + ox, cy + oy) for ox, oy in the equally spaced offset list `offs` (1/16 pel), through one grid call: the grid entry point takes
    centre +- k * step with step <= 16, so wider or centre-less sets ({-6,-2,2,6}, step 32) are read out of the smallest covering lattice"""
    offs = list(offs)
    if len(offs) == 1:
        c = cands.copy(); c['mvx'] = cx + offs[0]; c['mvy'] = cy + offs[0]
        return provider.errors(c).astype(np.int64).reshape(-1, 1, 1)
    d = offs[1] - offs[0]
    step = d
    while step > 16:
        step //= 2
    radius = -(-(offs[-1] - offs[0]) // (2 * step))
    shift = offs[0] + radius * step
    c = cands.copy(); c['mvx'] = cx + shift; c['mvy'] = cy + shift
    tab = provider.grid(c, step, radius).astype(np.int64)
    idx = [(o - offs[0]) // step for o in offs]
    return tab[:, idx, :][:, :, idx]


def _take_first_min_offsets(best_x, best_y, best_e, tab, cx, cy, offs, skip_zero):
    """loop order of the reference: y outer, x inner, strictly smaller wins"""
    for j, oy in enumerate(offs):
        for i, ox in enumerate(offs):
            if skip_zero and ox == 0 and oy == 0:
                continue
            e = tab[:, j, i]
            better = e < best_e
            best_e[better] = e[better]; best_x[better] = cx[better] + ox; best_y[better] = cy[better] + oy


def estimate_level(provider, width, height, block_size, previous=None, factor=2, double_res=False, bit_depth=10, unit_size=16, search_pattern=0):
    """One level of the MCTF motion search for the whole picture.  previous: None or (prev_x, prev_y) int arrays [prevH][prevW] of the coarser level.
    Returns dict(x, y, error, rmsme, overlap) with arrays [blocksY][blocksX] (vectors in 1/16 pel) -- MotionVector fields of MCTF.h:72-82."""
    bs = block_size
    bxn, byn = len(range(0, width - 7, bs)), len(range(0, height - 7, bs))       # `blockX + 8 <= origWidth` (:1174, :1388): partial blocks of >= 8 pels count
    gx, gy = np.meshgrid(np.arange(bxn) * bs, np.arange(byn) * bs)
    n = bxn * byn
    X = gx.reshape(-1).astype(np.int32); Y = gy.reshape(-1).astype(np.int32)
    W = (np.minimum(bs, width - X) & ~7).astype(np.uint16); H = (np.minimum(bs, height - Y) & ~7).astype(np.uint16)
    best_x = np.zeros(n, dtype=np.int32); best_y = np.zeros(n, dtype=np.int32); best_e = np.full(n, INT_MAX, dtype=np.int64)
    allb = np.arange(n)

    def cands_of(sel, mvx, mvy):
        c = np.zeros(len(sel), dtype=CAND_DT)
        c['x'] = X[sel]; c['y'] = Y[sel]; c['mvx'] = mvx; c['mvy'] = mvy; c['w'] = W[sel]; c['h'] = H[sel]
        return c

    # ---- stage A: predictors of the coarser level (3x3 neighbourhood, raster order) and the zero vector
    search_range = 8
    if previous is not None:
        search_range = 0 if double_res else (3 if search_pattern == 2 else 5)
        px, py = previous
        ph, pw = px.shape
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ty = Y // (2 * bs) + dy; tx = X // (2 * bs) + dx
                ok = (ty >= 0) & (ty < ph) & (tx >= 0) & (tx < pw)
                sel = allb[ok]
                if len(sel) == 0:
                    continue
                mvx = px[ty[ok], tx[ok]].astype(np.int32) * factor; mvy = py[ty[ok], tx[ok]].astype(np.int32) * factor
                e = provider.errors(cands_of(sel, mvx, mvy)).astype(np.int64)
                better = e < best_e[sel]
                idx = sel[better]
                best_e[idx] = e[better]; best_x[idx] = mvx[better]; best_y[idx] = mvy[better]
        e = provider.errors(cands_of(allb, 0, 0)).astype(np.int64)
        better = e < best_e
        best_e[better] = e[better]; best_x[better] = 0; best_y[better] = 0
    # ---- stage B: integer grid around trunc(prevBest / 16); search pattern 2 visits every second position of the first level (:1217)
    cx = _trunc_div16(best_x) * 16; cy = _trunc_div16(best_y) * 16
    d = 2 if (previous is None and search_pattern == 2) else 1
    offs = [16 * v for v in range(-search_range, search_range + 1, d)]
    tab = _offset_table(provider, cands_of(allb, 0, 0), cx, cy, offs)
    _take_first_min_offsets(best_x, best_y, best_e, tab, cx, cy, offs, False)
    # ---- stage C: sub-pel refinement around the running best (:1229-1287): +-12 step 4 (pattern 0), +-6 step 4 (1) or step 6 (2); then +-2 step 2; then +-1
    if double_res:
        rng = 12 if search_pattern == 0 else 6
        d1 = 6 if search_pattern == 2 else 4
        for offs in (list(range(-rng, rng + 1, d1)), [-2, 0, 2], [-1, 0, 1]):
            cx = best_x.copy(); cy = best_y.copy()
            tab = _offset_table(provider, cands_of(allb, 0, 0), cx, cy, offs)
            _take_first_min_offsets(best_x, best_y, best_e, tab, cx, cy, offs, True)
    # ---- stage D: final vectors of the upper and the left neighbour, along anti-diagonals
    bxi = (X // bs); byi = (Y // bs)
    for wave in range(1, bxn + byn - 1):
        sel = allb[(bxi + byi) == wave]
        up = sel[byi[sel] > 0]
        if len(up):
            src = up - bxn
            mvx = best_x[src].copy(); mvy = best_y[src].copy()
            e = provider.errors(cands_of(up, mvx, mvy)).astype(np.int64)
            better = e < best_e[up]
            idx = up[better]
            best_e[idx] = e[better]; best_x[idx] = mvx[better]; best_y[idx] = mvy[better]
        left = sel[bxi[sel] > 0]
        if len(left):
            src = left - 1
            mvx = best_x[src].copy(); mvy = best_y[src].copy()
            e = provider.errors(cands_of(left, mvx, mvy)).astype(np.int64)
            better = e < best_e[left]
            idx = left[better]
            best_e[idx] = e[better]; best_x[idx] = mvx[better]; best_y[idx] = mvy[better]
    rmsme = np.full(n, 0xffff, dtype=np.uint16); overlap = np.zeros(n, dtype=np.float64)
    err_out = best_e.copy()
    # ---- stage E: error scaling of the final level
    if double_res:
        var = provider.calc_var(cands_of(allb, 0, 0))
        bd_scale = float(1 << (2 * (10 - bit_depth)))
        wh = W.astype(np.float64) * H.astype(np.float64)
        dvar = var * bd_scale
        mse = best_e.astype(np.float64) * bd_scale / wh
        err_out = (20 * ((best_e.astype(np.float64) * bd_scale + 5.0) / (dvar + 5.0)) + mse / 50.0).astype(np.int64)      # (int) truncation
        rmsme = (0.5 + np.sqrt(mse)).astype(np.uint16)
        overlap = wh / float(unit_size * unit_size)
    shp = (byn, bxn)
    return dict(x=best_x.reshape(shp), y=best_y.reshape(shp), error=err_out.astype(np.int32).reshape(shp), rmsme=rmsme.reshape(shp), overlap=overlap.reshape(shp))


def subsample_luma(pic):
    """MCTF::subsampleLuma (MCTF.cpp:1072-1097): 2x2 average with rounding; pic is the unpadded picture [H][W]"""
    h, w = pic.shape[0] // 2, pic.shape[1] // 2
    p = pic[:2 * h, :2 * w].astype(np.int32)
    return ((p[0::2, 0::2] + p[1::2, 0::2] + p[0::2, 1::2] + p[1::2, 1::2] + 2) >> 2).astype(np.int16)


def pad_edge(pic, pad=128):
    """border replication as PelStorage::extendBorderPel (MCTF_PADDING = 128, CommonDef.h:520)"""
    return np.ascontiguousarray(np.pad(pic, pad, mode='edge'))


def estimate_pyramid(make_provider, org, ref, unit_size=16, add_level=False, bit_depth=10, search_pattern=0):
    """MCTF::motionEstimationMCTF (MCTF.cpp:666-724) for one neighbour picture: subsampled pyramids and four (five with add_level) chained levels.
    org / ref: unpadded pictures [H][W]; make_provider(org_pic, ref_pic) returns a provider for that pair of level pictures (it pads / uploads them).
    The motion-field arrays of the intermediate levels are sized as the reference sizes them (width / (unit * k) + 1): entries no block writes keep the
    default vector (0, 0).  Returns the final field dict of estimate_level with [ceil(H/unit)][ceil(W/unit)] entries."""
    H, W = org.shape
    o = [org]; r = [ref]
    for _ in range(3 if add_level else 2):
        o.append(subsample_luma(o[-1])); r.append(subsample_luma(r[-1]))

    def level(lv, bs, prev, factor, double_res, out_w, out_h):
        h, w = o[lv].shape
        f = estimate_level(make_provider(o[lv], r[lv]), w, h, bs, prev, factor, double_res, bit_depth, unit_size, search_pattern)
        fx = np.zeros((out_h, out_w), dtype=np.int32); fy = np.zeros((out_h, out_w), dtype=np.int32)
        fh, fw = min(out_h, f['x'].shape[0]), min(out_w, f['x'].shape[1])
        fx[:fh, :fw] = f['x'][:fh, :fw]; fy[:fh, :fw] = f['y'][:fh, :fw]
        return f, (fx, fy)

    u = unit_size
    prev = None
    if add_level:
        _, prev = level(3, 2 * u, None, 2, False, W // (u * 16) + 1, H // (u * 16) + 1)
    _, prev = level(2, 2 * u, prev, 2, False, W // (u * 8) + 1, H // (u * 8) + 1)
    _, prev = level(1, 2 * u, prev, 2, False, W // (u * 4) + 1, H // (u * 4) + 1)
    _, prev = level(0, 2 * u, prev, 2, False, W // (u * 2) + 1, H // (u * 2) + 1)
    final, _ = level(0, u, prev, 1, True, (W + u - 1) // u, (H + u - 1) // u)
    return final
