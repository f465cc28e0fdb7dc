"""Host-side candidate enumerators -- the control logic that stays on the CPU.

They restate WHICH integer positions the reference's motion search evaluates, so that a fixed candidate set can be
sent to the GPU as one batch and the decisions replayed on the host afterwards:

  * full_search_window : InterSearch::xSetSearchRange / xPatternSearch (EncoderLib/InterSearch.cpp:2183-2251)
  * tz_diamond_pattern : the static point pattern of xTZ8PointDiamondSearch (InterSearch.cpp:557-758) for every
                         distance 1, 2, 4 ... <= search range, plus the optional raster grid (:2491-2497)
"""
import numpy as np
from ._lib import MV_DT


def tz_diamond_points(dist, corners_at_dist1=False):
    """offsets (dx, dy) relative to the start point, in the evaluation order of xTZ8PointDiamondSearch"""
    d = dist
    if d == 1:
        if corners_at_dist1:       # InterSearch.cpp:576-620
            return [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
        return [(0, -1), (-1, 0), (1, 0), (0, 1)]
    if d <= 8:                     # :625-643: 4 axis points at d, 4 diagonal points at d/2
        h = d >> 1
        return [(0, -d), (-h, -h), (h, -h), (-d, 0), (d, 0), (-h, h), (h, h), (0, d)]
    pts = [(0, -d), (-d, 0), (d, 0), (0, d)]          # :688-705: 16 points on the diamond
    q = d >> 2
    for i in range(1, 4):
        pts += [(-q * i, -d + q * i), (q * i, -d + q * i), (-q * i, d - q * i), (q * i, d - q * i)]
    return pts


def tz_diamond_pattern(search_range, raster_step=0, include_centre=True, corners_at_dist1=False):
    """The fixed TZ candidate set around a start vector: centre, diamonds at d = 1,2,4,.. <= search_range and (optionally) the
    raster grid with the given step inside +-search_range (InterSearch.cpp:2491-2497).  Returns a MV_DT array."""
    pts = [(0, 0)] if include_centre else []
    d = 1
    while d <= search_range:
        pts += tz_diamond_points(d, corners_at_dist1)
        d <<= 1
    if raster_step:
        for y in range(-search_range, search_range + 1, raster_step):
            for x in range(-search_range, search_range + 1, raster_step):
                pts.append((x, y))
    out = np.zeros(len(pts), dtype=MV_DT)
    out['dx'] = [p[0] for p in pts]; out['dy'] = [p[1] for p in pts]
    return out


def full_search_window(pred_x, pred_y, search_range, pic_w, pic_h, x, y, w, h, margin):
    """SearchRange (left, right, top, bottom) in integer pels around an integer predictor, clipped so that every candidate
    block stays inside the padded picture (xClipMvSearch / xSetSearchRange, InterSearch.cpp:2134-2207)."""
    left = max(pred_x - search_range, -margin - x)
    right = min(pred_x + search_range, pic_w + margin - w - x)
    top = max(pred_y - search_range, -margin - y)
    bottom = min(pred_y + search_range, pic_h + margin - h - y)
    return left, right, top, bottom


def quad_order_grid(block, pic_w, pic_h):
    """Top-left positions of the block x block grid in quad-tree z-order: (x,y),(x+b,y),(x,y+b),(x+b,y+b) per 2x2 group, the way the encoder's
    partitioner visits them; an odd last block row / column is appended afterwards.  The dense search evaluates each quad on a shared window."""
    nbx, nby = pic_w // block, pic_h // block
    xs, ys = [], []
    for qy in range(0, nby - 1, 2):
        for qx in range(0, nbx - 1, 2):
            for (dx, dy) in ((0, 0), (1, 0), (0, 1), (1, 1)):
                xs.append((qx + dx) * block); ys.append((qy + dy) * block)
    if nbx % 2:
        for by in range(nby - nby % 2):
            xs.append((nbx - 1) * block); ys.append(by * block)
    if nby % 2:
        for bx in range(nbx):
            xs.append(bx * block); ys.append((nby - 1) * block)
    return np.array(xs, dtype=np.int32), np.array(ys, dtype=np.int32)


def pyramid_lists(base, levels, pic_w, pic_h):
    """Block lists for vvb_sad_search_pyramid: level l holds blocks of size base << l; block j of level l+1 is the parent of blocks
    4j..4j+3 of level l (z-order).  Blocks that no larger block covers (picture size not a multiple of the larger size) are appended
    after the children of the level above, together with their own descendants.  Returns [(xs, ys)] per level."""
    lists = [[] for _ in range(levels)]

    def emit(level, x, y):
        lists[level].append((x, y))
        if level > 0:
            s = base << (level - 1)
            for (dx, dy) in ((0, 0), (1, 0), (0, 1), (1, 1)):
                emit(level - 1, x + dx * s, y + dy * s)

    top = base << (levels - 1)
    for y in range(0, pic_h - top + 1, top):
        for x in range(0, pic_w - top + 1, top):
            emit(levels - 1, x, y)
    for l in range(levels - 2, -1, -1):
        s = base << l
        wc, hc = (pic_w // (2 * s)) * 2 * s, (pic_h // (2 * s)) * 2 * s      # area covered by the level above
        for y in range(0, (pic_h // s) * s, s):
            for x in range(0, (pic_w // s) * s, s):
                if x >= wc or y >= hc:
                    emit(l, x, y)
    return [(np.array([p[0] for p in L], dtype=np.int32), np.array([p[1] for p in L], dtype=np.int32)) for L in lists]


# ---- fractional refinement control (InterSearch::xPatternSearchFracDIF, EncoderLib/InterSearch.cpp:2683-2725, with m_fastSubPel = 0) --------------
# visiting order of the two rounds of xPatternRefinement (:67-91): offsets in half-pel / quarter-pel units
REFINE_HALF = ((0, 0), (0, -1), (0, 1), (-1, 0), (1, 0), (-1, -1), (1, -1), (-1, 1), (1, 1))
REFINE_QUARTER = ((0, 0), (0, -1), (0, 1), (-1, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (1, 1))


def subpel_refinement(table, mv_int, mv_cost, quarter_round=True):
    """Replays the half-pel and the quarter-pel round on the 7x7 distortion table of vvb_frac_cost_grid (table[j+3][i+3], quarter-pel offsets) for one
    block.  mv_int: integer vector (x, y); mv_cost(x, y, cost_scale) -> rate of the vector in the units of that round (cost scale 1: half pel, 0: quarter
    pel; RdCost::getCostOfVectorWithPredictor).  Every round starts from MAX_DISTORTION and takes the first strictly smaller cost in visiting order.
    quarter_round=False stops after the half-pel round (AMVR half-pel mode, the only mode that uses the alternative half-pel filter, :2712).
    Returns (half_offset, quarter_offset, cost): the offsets rcMvHalf / rcMvQter of the reference and the final cost."""
    best = None; half = (0, 0)
    bx, by = mv_int[0] << 1, mv_int[1] << 1                    # rcMvHalf base = rcMvInt << 1
    for (dx, dy) in REFINE_HALF:
        c = int(table[2 * dy + 3][2 * dx + 3]) + mv_cost(bx + dx, by + dy, 1)
        if best is None or c < best:
            best = c; half = (dx, dy)
    if not quarter_round:
        return half, (0, 0), best
    qbx, qby = (bx + half[0]) << 1, (by + half[1]) << 1       # rcMvQter base = ((rcMvInt << 1) + rcMvHalf) << 1
    best = None; quarter = (0, 0)
    for (dx, dy) in REFINE_QUARTER:
        c = int(table[2 * half[1] + dy + 3][2 * half[0] + dx + 3]) + mv_cost(qbx + dx, qby + dy, 0)
        if best is None or c < best:
            best = c; quarter = (dx, dy)
    return half, quarter, best


# ---- m_fastSubPel = 1 (presets fast ... slow: vvencCfg.cpp:2718-2944): xPatternRefinement visits a subset of the nine positions -------------------------
# The half-pel round stops early (InterSearch.cpp:808-811) and classifies the cost surface into a pattern id (:886-969); the quarter-pel round then only
# visits the positions that pattern allows (s_skipQpelPosition, :93-137; below as one 9-bit mask per pattern, bit i = position i is skipped) and keeps the
# half-pel round's best cost as its starting threshold (:769).  Every visited position is still the two-pass interpolation of the 7x7 table.
SKIP_QPEL_MASK = (510, 479, 447, 509, 469, 429, 507, 347, 187, 123, 479, 347, 447, 187, 485, 479, 469, 447, 429, 175, 509, 429, 507, 187, 343, 509, 469, 507, 347,
                  447, 507, 187, 479, 507, 347, 447, 509, 429, 479, 509, 469, 0)
_U64 = (1 << 64) - 1
MAX_DISTORTION = _U64                 # std::numeric_limits<Distortion>::max(), CommonDef.h:202


def _pattern_id_after_half_round(dist_h, best_dir, pattern_id):
    """the switch of InterSearch.cpp:886-969 in uint64 wrap-around arithmetic (Distortion is uint64_t; positions not visited hold MAX_DISTORTION)"""
    TH, TL, SH = 17, 15, 4
    d = list(dist_h)

    def ratio(a, b, hi, lo):          # distH[a] <<= shift; > TH * distH[b] ? hi : ( < TL * distH[b] ? lo : 0 )
        d[a] = (d[a] << SH) & _U64
        return hi if d[a] > ((TH * d[b]) & _U64) else (lo if d[a] < ((TL * d[b]) & _U64) else 0)

    def slope(a, c, b):               # distH[a] - distH[c] > distH[c] - distH[b]
        return ((d[a] - d[c]) & _U64) > ((d[c] - d[b]) & _U64)

    p = pattern_id
    if best_dir == 0:
        p += ratio(3, 4, 2, 1)
        p += ratio(1, 2, 6, 3)
    elif best_dir == 1:
        p += ratio(5, 6, 4, 2); p += 1 if slope(2, 0, 1) else 0; p += 0 if p == 41 else 8
    elif best_dir == 2:
        p += ratio(7, 8, 4, 2); p += 1 if slope(1, 0, 2) else 0; p += 0 if p == 41 else 13
    elif best_dir == 3:
        p += 1 if slope(4, 0, 3) else 0; p += ratio(5, 7, 4, 2); p += 0 if p == 41 else 18
    elif best_dir == 4:
        p += 1 if slope(3, 0, 4) else 0; p += ratio(6, 8, 4, 2); p += 0 if p == 41 else 23
    elif best_dir == 5:
        p += 1 if slope(6, 1, 5) else 0; p += 2 if slope(7, 3, 5) else 0; p += 0 if p == 41 else 28
    elif best_dir == 6:
        p += 1 if slope(5, 1, 6) else 0; p += 2 if slope(8, 4, 6) else 0; p += 0 if p == 41 else 31
    elif best_dir == 7:
        p += 1 if slope(8, 2, 7) else 0; p += 2 if slope(5, 3, 7) else 0; p += 0 if p == 41 else 34
    elif best_dir == 8:
        p += 1 if slope(7, 2, 8) else 0; p += 2 if slope(6, 4, 8) else 0; p += 0 if p == 41 else 37
    return p


def subpel_refinement_fast(table, mv_int, mv_cost, quarter_round=True):
    """InterSearch::xPatternSearchFracDIF with m_fastSubPel = 1 replayed on the 7x7 table of vvb_frac_cost_grid.  Same arguments as subpel_refinement.
    Returns (half_offset, quarter_offset or None, cost): quarter_offset is None when the quarter-pel round does not run (pattern id 0 or AMVR half-pel mode) --
    the reference then leaves rcMvQter untouched."""
    bx, by = mv_int[0] << 1, mv_int[1] << 1
    best = MAX_DISTORTION; best_dir = 0
    dist_h = [MAX_DISTORTION] * 9
    pattern = 41
    for i, (dx, dy) in enumerate(REFINE_HALF):
        if (SKIP_QPEL_MASK[pattern] >> i) & 1:
            continue
        if (i == 5 and best_dir == 0) or (i == 7 and best_dir == 1) or (i == 8 and best_dir in (1, 3, 5)):
            break
        c = int(table[2 * dy + 3][2 * dx + 3]) + mv_cost(bx + dx, by + dy, 1)
        dist_h[i] = c
        if c < best:
            best = c; best_dir = i
    half = REFINE_HALF[best_dir]
    pattern = _pattern_id_after_half_round(dist_h, best_dir, pattern) - 41
    if not quarter_round or pattern == 0:
        return half, None, best
    qbx, qby = (bx + half[0]) << 1, (by + half[1]) << 1
    best_dir = 0
    for i, (dx, dy) in enumerate(REFINE_QUARTER):
        if (SKIP_QPEL_MASK[pattern] >> i) & 1:
            continue
        c = int(table[2 * half[1] + dy + 3][2 * half[0] + dx + 3]) + mv_cost(qbx + dx, qby + dy, 0)
        if c < best:                  # the threshold is the half-pel round's best (:769), cost scales differ on purpose
            best = c; best_dir = i
    return half, REFINE_QUARTER[best_dir], best
