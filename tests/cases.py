"""Seeded parity cases shared by the golden-vector generator (tests/golden/make_golden.py), the oracle tests and
the GPU parity tests.  Inputs come from numpy's frozen legacy MT19937 stream (np.random.RandomState), so a case
is fully described by its parameter row + seed and the fixtures only need to carry the expected outputs.

The matrix mirrors the reference's differential unit tests (test/vvenc_unit_test/vvenc_unit_test.cpp:1885-2180
RdCost, :912-1211 TCoeffOps, :1441-1560 MCTF, :2182-2266 affine) and adds what they do not pin (DF_SSE*, real
DCT/DST matrices through xT, QuantCore, needRdoq, Sobel) -- SURVEY.md section 8c.
"""
import numpy as np

FAM_SSE, FAM_SAD, FAM_HAD, FAM_HAD_FAST, FAM_HAD_2SAD = range(5)


def aligned(shape, dtype, align=64):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


def pel_block(rs, h, stride, kind, bit_depth=10):
    hi = (1 << bit_depth)
    if kind == 0:
        a = rs.randint(0, hi, size=(h, stride))
    elif kind == 1:
        a = np.full((h, stride), hi - 1)
    elif kind == 2:
        a = np.zeros((h, stride))
    else:   # smooth-ish: small differences, exercises small SATD values / rounding
        a = 512 + rs.randint(-6, 7, size=(h, stride))
    out = aligned((h, stride), np.int16)
    out[:] = a
    return out


def dist_cases():
    """rows: family, w, h, org_stride, cur_stride, subShift, kind_org, kind_cur, bit_depth, seed"""
    rows = []
    rs = np.random.RandomState(1234)
    seed = 1000
    widths = [1, 2, 4, 8, 16, 32, 64, 128]
    heights = [1, 2, 4, 6, 8, 12, 16, 24, 32, 64, 128]
    for w in widths:
        for h in heights:
            for fam in range(5):
                if fam >= FAM_HAD and (w < 2 or h % 2):
                    continue
                if fam == FAM_SSE and w == 1:
                    continue          # RdCost::getDistPart routes w==1 to scalar xGetSSE (RdCost.cpp:275-278)
                if fam == FAM_HAD_2SAD and (w < 4 or h % 4):
                    continue          # RdCost.cpp:1783-1784 walks h/4 rows of 4*w pels; SIMD loads 16 at a time (RdCostX86.h:2585)
                for rep in range(2 if w * h <= 1024 else 1):
                    compact = fam == FAM_HAD_2SAD      # RdCost.cpp:1778 CHECKD: compact, aligned buffers
                    so = w if compact else w + int(rs.randint(0, 40))
                    sc = w if compact else w + int(rs.randint(0, 40))
                    ss = int(rs.randint(0, 2)) if (fam == FAM_SAD and h >= 2 and h % 2 == 0) else 0
                    kinds = [(0, 0), (1, 2), (3, 3), (0, 3)][(seed + rep) % 4]
                    bd = 8 if (seed % 7 == 0) else 10
                    rows.append([fam, w, h, so, sc, ss, kinds[0], kinds[1], bd, seed])
                    seed += 1
    return np.array(rows, dtype=np.int32)


def dist_inputs(row):
    fam, w, h, so, sc, ss, ko, kc, bd, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    return pel_block(rs, h, so, ko, bd), pel_block(rs, h, sc, kc, bd)


def tq_cases():
    """rows: trHor, trVer, w, h, stride, amp, qp, isIRAP, bit_depth, seed   (tr: 0 DCT2, 1 DCT8, 2 DST7)"""
    rows = []
    seed = 5000
    rs = np.random.RandomState(99)
    for w in (4, 8, 16, 32, 64):
        for h in (4, 8, 16, 32, 64):
            for (th, tv) in ((0, 0), (2, 2), (1, 2), (2, 1), (1, 1)):
                if (th or tv) and (w > 32 or h > 32):
                    continue
                for amp in (1023, 200, 12):
                    bd = 8 if seed % 5 == 0 else 10
                    a = min(amp, (1 << bd) - 1)
                    rows.append([th, tv, w, h, w + int(rs.randint(0, 6)), a, int(rs.randint(8, 52)), int(rs.randint(0, 2)), bd, seed])
                    seed += 1
    return np.array(rows, dtype=np.int32)


def tq_inputs(row):
    th, tv, w, h, st, amp, qp, irap, bd, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    if seed % 11 == 0:
        a = np.full((h, st), amp)
    elif seed % 11 == 1:
        a = np.where(rs.randint(0, 2, size=(h, st)) > 0, amp, -amp)
    else:
        a = rs.randint(-amp, amp + 1, size=(h, st))
    return a.astype(np.int16)


def itq_cases():
    """inverse path rows: trHor, trVer, w, h, stride (of the residual written), kind, qp, bit_depth, seed
    kind 0: sparse decaying levels (what a quantiser emits), 1: dense +-amp levels everywhere (zero-out region included: the
    reference ignores it), 2: extreme levels (int16 limits: dequant input/output clipping), 3: DC only, 4: all zero"""
    rows = []
    seed = 15000
    rs = np.random.RandomState(1234)
    for w in (4, 8, 16, 32, 64):
        for h in (4, 8, 16, 32, 64):
            for (th, tv) in ((0, 0), (2, 2), (1, 2), (2, 1), (1, 1)):
                if (th or tv) and (w > 32 or h > 32):
                    continue
                for kind in (0, 1, 2):
                    bd = 8 if seed % 5 == 0 else 10
                    qp = int(rs.randint(0, 64)) if kind else int(rs.randint(12, 48))
                    rows.append([th, tv, w, h, w + int(rs.randint(0, 6)), kind, qp, bd, seed])
                    seed += 1
            rows.append([0, 0, w, h, w, 3, int(rs.randint(0, 64)), 10, seed]); seed += 1
    rows.append([0, 0, 16, 16, 16, 4, 30, 10, seed])
    return np.array(rows, dtype=np.int32)


def itq_inputs(row):
    th, tv, w, h, st, kind, qp, bd, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    q = np.zeros((h, w), dtype=np.int64)
    if kind == 0:
        yy, xx = np.mgrid[0:h, 0:w]
        mag = 60.0 / (1.0 + 0.9 * (xx + yy))
        q = np.rint(rs.standard_normal((h, w)) * mag).astype(np.int64)
        q[rs.rand(h, w) < 0.35] = 0
    elif kind == 1:
        amp = int(rs.choice([1, 7, 300, 4000]))
        q = rs.randint(-amp, amp + 1, size=(h, w))
    elif kind == 2:
        q = rs.choice(np.array([-32768, -32767, -20000, -1, 0, 1, 12345, 32767]), size=(h, w))
    elif kind == 3:
        q[0, 0] = int(rs.randint(-2000, 2000))
    return np.ascontiguousarray(q.astype(np.int16))


def rt_cases():
    """TU round-trip rows: trHor, trVer, w, h, org_stride, pred_stride, amp, qp, isIRAP, bit_depth, seed"""
    rows = []
    seed = 17000
    rs = np.random.RandomState(4321)
    for w in (4, 8, 16, 32, 64):
        for h in (4, 8, 16, 32, 64):
            for (th, tv) in ((0, 0), (2, 2), (1, 2)):
                if (th or tv) and (w > 32 or h > 32):
                    continue
                for amp in (1023, 60, 3):
                    bd = 8 if seed % 7 == 0 else 10
                    rows.append([th, tv, w, h, w + int(rs.randint(0, 9)), w + int(rs.randint(0, 9)), min(amp, (1 << bd) - 1),
                                 int(rs.randint(10, 50)), int(rs.randint(0, 2)), bd, seed])
                    seed += 1
    return np.array(rows, dtype=np.int32)


def rt_inputs(row):
    th, tv, w, h, so, ps, amp, qp, irap, bd, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    mx = (1 << bd) - 1
    org = rs.randint(0, mx + 1, size=(h, so))
    base = org[:, :w] if amp < mx else rs.randint(0, mx + 1, size=(h, w))
    pred = np.zeros((h, ps), dtype=np.int64)
    pred[:, :w] = np.clip(base + rs.randint(-amp, amp + 1, size=(h, w)), 0, mx)
    return np.ascontiguousarray(org.astype(np.int16)), np.ascontiguousarray(pred.astype(np.int16))


def mctf_cases():
    """rows: w, h, mvx, mvy (1/16 pel), tap4, bit_depth, seed"""
    rows = []
    seed = 9000
    rs = np.random.RandomState(7)
    for (w, h) in ((8, 8), (16, 16), (16, 8), (32, 32), (64, 64), (24, 40), (8, 64)):
        rows.append([w, h, 16 * int(rs.randint(-3, 4)), 16 * int(rs.randint(-3, 4)), 0, 10, seed]); seed += 1
        for tap4 in (0, 1):
            for k in range(10):
                mvx = int(rs.randint(-64, 64)); mvy = int(rs.randint(-64, 64))
                if (mvx | mvy) & 15 == 0:
                    mvx += 5
                rows.append([w, h, mvx, mvy, tap4, 8 if k == 9 else 10, seed]); seed += 1
    return np.array(rows, dtype=np.int32)


MCTF_MARGIN = 12


def mctf_inputs(row):
    w, h, mvx, mvy, tap4, bd, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    m = MCTF_MARGIN
    org = rs.randint(0, 1 << bd, size=(h, w + 5)).astype(np.int16)
    buf = rs.randint(0, 1 << bd, size=(h + 2 * m, w + 2 * m)).astype(np.int16)
    return org, buf


def affine_cases():
    """rows: w, h, pred_stride, deriv_stride, six_param, seed"""
    rows = []
    seed = 12000
    rs = np.random.RandomState(5)
    for (w, h) in ((16, 16), (32, 16), (16, 32), (64, 64), (128, 64), (128, 128)):
        for six in (0, 1):
            rows.append([w, h, w + int(rs.randint(0, 9)), w + int(rs.randint(0, 9)), six, seed]); seed += 1
    return np.array(rows, dtype=np.int32)


def affine_inputs(row):
    w, h, ps, ds, six, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    pred = rs.randint(0, 1024, size=(h, ps)).astype(np.int16)
    resi = rs.randint(-1023, 1024, size=(h, ps)).astype(np.int16)
    if seed % 3 == 0:      # extremes, as the reference's MinMaxGenerator does (vvenc_unit_test.cpp:170-202)
        gx = np.where(rs.randint(0, 2, size=(h, ds)) > 0, 4095, -4096).astype(np.int16)
        gy = np.where(rs.randint(0, 2, size=(h, ds)) > 0, 4095, -4096).astype(np.int16)
    else:
        gx = rs.randint(-4096, 4096, size=(h, ds)).astype(np.int16)
        gy = rs.randint(-4096, 4096, size=(h, ds)).astype(np.int16)
    return pred, resi, gx, gy


def search_case(seed=777, W=192, H=128, margin=48):
    """one small picture pair + block list for the full-search replay; returns dict"""
    rs = np.random.RandomState(seed)
    S = W + 2 * margin
    base = rs.randint(0, 1024, size=(H + 2 * margin + 8, S + 8))
    # low-pass so that argmins are non-trivial, then pan by (3,-2) plus noise
    sm = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4
    org = sm[4:4 + H + 2 * margin, 4:4 + S].astype(np.int16)
    ref = np.clip(sm[4 + 3:4 + 3 + H + 2 * margin, 4 - 2:4 - 2 + S] + rs.randint(-9, 10, size=org.shape), 0, 1023).astype(np.int16)
    blks = []
    for (bw, bh) in ((4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 16), (32, 16), (64, 32), (128, 64)):
        for k in range(3):
            x = int(rs.randint(0, (W - bw) // 4 + 1)) * 4; y = int(rs.randint(0, (H - bh) // 4 + 1)) * 4
            r = int(rs.randint(2, 17))
            blks.append([x, y, bw, bh, -r, r, -r, r, int(rs.randint(-60, 60)), int(rs.randint(-60, 60))])
    return dict(org=np.ascontiguousarray(org), ref=np.ascontiguousarray(ref), stride=S, margin=margin, W=W, H=H,
                blk=np.array(blks, dtype=np.int32), lam=57.25, cost_scale=2, imv_shift=0)


def mctf_apply_case(seed, W=96, H=64, margin=24, num_refs=4, bs=16, bit_depth=10):
    """one small picture + num_refs neighbour pictures (noisy, shifted copies) and per-block motion vectors with error / rmsme, for the MCTF
    apply stage (xFinalizeBlkLine).  Returns dict(org, refs[list], stride, margin, W, H, mvs[num_refs][blocks][4], strengths, ws, sigma)"""
    rs = np.random.RandomState(seed)
    mx = (1 << bit_depth) - 1
    S = W + 2 * margin
    base = rs.randint(0, mx + 1, size=(H + 2 * margin + 8, S + 8))
    sm = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4
    org = np.ascontiguousarray(sm[4:4 + H + 2 * margin, 4:4 + S].astype(np.int16))
    refs = []
    for r in range(num_refs):
        sh = (int(rs.randint(-2, 3)), int(rs.randint(-2, 3)))
        amp = [3, 12, 60, mx][r % 4]
        ref = np.clip(sm[4 + sh[0]:4 + sh[0] + H + 2 * margin, 4 + sh[1]:4 + sh[1] + S] + rs.randint(-amp, amp + 1, size=org.shape), 0, mx).astype(np.int16)
        refs.append(np.ascontiguousarray(ref))
    nb = ((W + bs - 1) // bs) * ((H + bs - 1) // bs)
    mvs = np.zeros((num_refs, nb, 4), dtype=np.int32)
    mvs[..., 0] = rs.randint(-16 * 6, 16 * 6 + 1, size=(num_refs, nb)); mvs[..., 1] = rs.randint(-16 * 6, 16 * 6 + 1, size=(num_refs, nb))
    mvs[..., 2] = rs.choice([3, 20, 49, 50, 75, 100, 101, 400], size=(num_refs, nb))
    mvs[..., 3] = rs.choice([0, 1, 5, 22, 23, 60], size=(num_refs, nb))
    mvs[0, 0, :2] = 0
    strengths = np.array([0.85, 0.57, 0.41, 0.33, 0.30, 0.20, 0.18, 0.15][:num_refs], dtype=np.float64)
    return dict(org=org, refs=refs, stride=S, margin=margin, W=W, H=H, mvs=mvs, strengths=strengths, ws=0.4 * float(rs.choice([0.5, 1.0, 1.5])),
                sigma=float(rs.choice([9 * (128.0 + 3.0 / 256.0 * q * q * q) for q in (22, 32, 42)])) / (1.0 if bit_depth == 10 else 16.0), bs=bs, bd=bit_depth,
                num_refs=num_refs)


MCTF_APPLY_CASES = ((31, 96, 64, 4, 16, 10, 0, 1), (32, 64, 48, 8, 8, 10, 0, 1), (33, 128, 64, 6, 32, 10, 0, 1), (34, 96, 40, 3, 16, 8, 0, 1),
                    (35, 96, 64, 5, 16, 10, 1, 1), (36, 72, 56, 2, 16, 10, 0, 0))      # seed, W, H, refs, bs, bit depth, tap4, planar


def frac_case(seed=4242, W=160, H=96, margin=24, bit_depth=10):
    """picture pair + block lists for the fractional-pel refinement grid: list of (family, w, h, blocks[n][6] = x, y, w, h, mvx, mvy)"""
    rs = np.random.RandomState(seed)
    mx = (1 << bit_depth) - 1
    S = W + 2 * margin
    base = rs.randint(0, mx + 1, size=(H + 2 * margin + 8, S + 8))
    sm = (base + np.roll(base, 1, 0) + np.roll(base, 1, 1) + np.roll(base, (1, 1), (0, 1))) // 4
    org = np.ascontiguousarray(sm[4:4 + H + 2 * margin, 4:4 + S].astype(np.int16))
    ref = np.ascontiguousarray(np.clip(sm[4 + 1:4 + 1 + H + 2 * margin, 4 - 2:4 - 2 + S] + rs.randint(-9, 10, size=org.shape), 0, mx).astype(np.int16))
    if seed % 2:                       # extreme content: full-range checkerboard noise in a corner exercises the intermediate range
        ref[margin:margin + 40, margin:margin + 40] = np.where(rs.randint(0, 2, size=(40, 40)) > 0, mx, 0)
    lists = []
    for (fam, w, h) in ((1, 8, 8), (2, 8, 8), (1, 16, 16), (2, 16, 16), (2, 32, 32), (2, 64, 64), (1, 16, 8), (1, 32, 16), (1, 8, 32), (1, 64, 64)):
        n = 6 if w * h <= 256 else 3
        b = np.zeros((n, 6), dtype=np.int32)
        for k in range(n):
            b[k] = (int(rs.randint(0, W - w + 1)), int(rs.randint(0, H - h + 1)), w, h, int(rs.randint(-9, 10)), int(rs.randint(-9, 10)))
        b[0, :2] = 0
        lists.append((fam, w, h, b))
    return dict(org=org, ref=ref, stride=S, margin=margin, W=W, H=H, bd=bit_depth, lists=lists)


FRAC_CASES = ((4242, 10), (4243, 10), (4244, 8))


def frac_filter_of(list_index):
    """(reduce_tap, alt_hpel) used for block list `list_index` of a frac_case: cycles through the filter sets (ReduceFilterME 2 first: what the presets use)"""
    return ((2, 0), (2, 0), (0, 0), (1, 0), (2, 1), (0, 1), (2, 0), (1, 1), (0, 0), (2, 0))[list_index % 10]


# ---- dependent quantisation (DepQuant::xQuantDQ) ------------------------------------------------------------------------------------------------------------
def dq_cases():
    """rows: w, h, bit_depth, qp, lambda * 1000, scale, decay * 10, mtsIdx (0 or 2..5), lfnstIdx, sbtInfo, intraCu, ctxInitId, seed
    The CABAC contexts are initialised for slice QP = qp and init type ctxInitId, as a slice start does."""
    rows = []
    rs = np.random.RandomState(4711)
    seed = 9000
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 16), (32, 8), (16, 64), (64, 32), (32, 16), (4, 32), (64, 4), (16, 8)]:
        for k in range(14):
            bd = 8 if k % 5 == 4 else 10
            qp = int(rs.choice([17, 22, 27, 32, 37, 42, 51]))
            lam = float(rs.choice([3.0, 11.7, 30.0, 57.3, 120.0, 800.0, 4000.0]))
            scale = int(rs.choice([5, 20, 60, 200, 600, 2000, 30000]))
            decay = int(rs.choice([1, 5, 10, 15]))
            mts = int(rs.choice([0, 0, 2, 3, 5])) if (w <= 32 and h <= 32) else 0
            lf = int(rs.choice([0, 0, 0, 1, 2])) if mts == 0 else 0
            sbt = int(rs.choice([0, 0, 0, 1])) if (mts == 0 and lf == 0) else 0
            intra = 1 if lf else (0 if sbt else int(rs.randint(2)))
            rows.append([w, h, bd, qp, int(lam * 1000), scale, decay, mts, lf, sbt, intra, k % 3, seed])
            seed += 1
    # levels above 127 (large coefficients, low QP, flat spectrum): here the reference's scalar and x86 state updates part ways
    for (w, h) in [(16, 16), (32, 32), (64, 64), (32, 16), (8, 8)]:
        for (qp, lam) in ((27, 1500.0), (22, 4000.0), (17, 800.0), (27, 4000.0)):
            rows.append([w, h, 10, qp, int(lam * 1000), 30000, 1, 0, 0, 0, seed & 1, seed % 3, seed])
            seed += 1
    return np.array(rows, dtype=np.int64)


def dq_inputs(row):
    w, h, bd, qp, lam1000, scale, decay10, mts, lf, sbt, intra, init_id, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    coef = rs.laplace(0, scale, size=(h, w)) * (1.0 / (1 + np.add.outer(np.arange(h), np.arange(w))) ** (decay10 / 10.0))
    coef = np.clip(coef, -32768, 32767).astype(np.int32)
    if w > 32:
        coef[:, 32:] = 0          # what the transform's zero-out leaves
    if h > 32:
        coef[32:, :] = 0
    return coef


def dq_zero_out(row):
    w, h, mts, sbt = int(row[0]), int(row[1]), int(row[7]), int(row[9])
    return 1 if (mts > 1 or (sbt and w <= 32 and h <= 32)) else 0       # DepQuant.cpp:1155


# ---- transform skip / chroma components (TrQuant::xTransformSkip, Quant::quant with cQP.per / rem( true ), xNeedRDOQ's chroma constant) -------------------------
def ts_cases():
    """rows: w, h, stride, bit_depth, amp, qp, isIRAP, signHiding, depQuant, transformSkip, inputDelta, comp (0 luma, 1 Cb), seed"""
    rows = []
    rs = np.random.RandomState(2024)
    seed = 12000
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (4, 8), (16, 4), (32, 8), (8, 32), (16, 32), (64, 64), (64, 16)]:
        for bd in (8, 10):
            for k in range(10):
                ts = 1 if (k % 2 == 0 and w <= 32 and h <= 32) else 0
                amp = int(rs.choice([3, 12, 60, 300, (1 << bd) - 1]))
                rows.append([w, h, w + int(rs.randint(0, 5)), bd, amp, int(rs.randint(-6 * (bd - 8), 64)), int(rs.randint(2)), int(rs.randint(2)), int(rs.randint(2)), ts,
                             int(rs.choice([0, 0, 2])) if bd == 10 else 0, int(rs.randint(2)), seed])
                seed += 1
    return np.array(rows, dtype=np.int32)


def ts_inputs(row):
    w, h, st, bd, amp = [int(v) for v in row[:5]]
    rs = np.random.RandomState(int(row[12]))
    return rs.randint(-amp, amp + 1, size=(h, st)).astype(np.int16)


def dq_chroma_cases():
    """chroma components: rows w, h, bit_depth, qp, lambda * 1000, scale, decay * 10, lfnstIdx (0), intraCu, ctxInitId, seed"""
    rows = []
    rs = np.random.RandomState(815)
    seed = 15000
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 16), (32, 8), (16, 32), (64, 64)]:
        for k in range(8):
            bd = 8 if k % 4 == 3 else 10
            rows.append([w, h, bd, int(rs.choice([17, 27, 37, 47])), int(float(rs.choice([3.0, 11.7, 57.3, 800.0])) * 1000), int(rs.choice([5, 20, 200, 2000, 30000])),
                         int(rs.choice([1, 5, 10])), 0, int(rs.randint(2)), k % 3, seed])
            seed += 1
    return np.array(rows, dtype=np.int64)


def dq_chroma_inputs(row):
    w, h, bd, qp, lam1000, scale, decay10 = [int(v) for v in row[:7]]
    rs = np.random.RandomState(int(row[10]))
    coef = rs.laplace(0, scale, size=(h, w)) * (1.0 / (1 + np.add.outer(np.arange(h), np.arange(w))) ** (decay10 / 10.0))
    coef = np.clip(coef, -32768, 32767).astype(np.int32)
    if w > 32:
        coef[:, 32:] = 0
    if h > 32:
        coef[32:, :] = 0
    return coef


# ---- fast RDOQ (QuantRDOQ2::xRateDistOptQuant, m_RDOQ == 2) -----------------------------------------------------------------------------------------------------
def rdoq_cases():
    """rows: w, h, bit_depth, qp, lambda * 1000, scale, decay * 10, comp (0 Y, 1 Cb, 2 Cr), lfnstIdx, sbtInfo, intraCu, signHiding, cbCbf, thrVal, ctxInitId, seed
    The CABAC contexts are initialised for slice QP = qp and init type ctxInitId, as a slice start does."""
    rows = []
    rs = np.random.RandomState(2602)
    seed = 21000
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (8, 4), (4, 16), (32, 8), (16, 64), (64, 32), (32, 16), (4, 32), (64, 4), (16, 8)]:
        for k in range(16):
            bd = 8 if k % 5 == 4 else 10
            qp = int(rs.choice([17, 22, 27, 32, 37, 42, 51]))
            lam = float(rs.choice([3.0, 11.7, 30.0, 57.3, 120.0, 800.0, 4000.0]))
            scale = int(rs.choice([5, 20, 60, 200, 600, 2000, 30000]))
            decay = int(rs.choice([1, 5, 10, 15]))
            comp = int(rs.choice([0, 0, 0, 1, 2]))
            lf = int(rs.choice([0, 0, 0, 1, 2]))
            sbt = int(rs.choice([0, 0, 0, 1])) if (lf == 0 and comp == 0) else 0
            intra = 1 if lf else (0 if sbt else int(rs.randint(2)))
            rows.append([w, h, bd, qp, int(lam * 1000), scale, decay, comp, lf, sbt, intra, k & 1, int(rs.randint(2)) if comp == 2 else 0, int(rs.choice([8, 8, 4, 16])), k % 3, seed])
            seed += 1
    return np.array(rows, dtype=np.int64)


def rdoq_inputs(row):
    w, h, bd, qp, lam1000, scale, decay10 = [int(v) for v in row[:7]]
    rs = np.random.RandomState(int(row[15]))
    coef = rs.laplace(0, scale, size=(h, w)) * (1.0 / (1 + np.add.outer(np.arange(h), np.arange(w))) ** (decay10 / 10.0))
    coef = np.clip(coef, -32768, 32767).astype(np.int32)
    if w > 32:
        coef[:, 32:] = 0          # what the transform's zero-out leaves
    if h > 32:
        coef[32:, :] = 0
    return coef


def rdoq_ts_cases():
    """transform-skipped TUs (QuantRDOQ::rateDistOptQuantTS): rows w, h, bit_depth, qp, lambda * 1000, amp, kind (0 uniform, 1 laplace, 2 sparse), comp (0 Y, 1 Cb), intraCu, inputDelta, ctxInitId, seed"""
    rows = []
    rs = np.random.RandomState(3107)
    seed = 23000
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 16), (32, 8), (16, 32), (4, 32), (16, 4)]:
        for k in range(14):
            bd = 8 if k % 5 == 4 else 10
            rows.append([w, h, bd, int(rs.choice([2, 17, 22, 27, 32, 37, 42, 51])), int(float(rs.choice([3.0, 11.7, 30.0, 57.3, 120.0, 800.0, 4000.0])) * 1000),
                         int(rs.choice([2, 6, 20, 60, 200, 1023])), k % 3, int(rs.randint(2)), int(rs.randint(2)), int(rs.choice([0, 0, 2])) if bd == 10 else 0, k % 3, seed])
            seed += 1
    return np.array(rows, dtype=np.int64)


def rdoq_ts_inputs(row):
    """TrQuant::xTransformSkip copies the residual unscaled (TrQuant.cpp:1050-1064), so the quantiser sees values inside the bit depth; every second row feeds values
    scaled up by the transform shift instead, so that large levels and the exhaustion of the context-coded bins are covered as well"""
    w, h, bd, qp, lam1000, amp, kind = [int(v) for v in row[:7]]
    rs = np.random.RandomState(int(row[11]))
    if kind == 0:
        resi = rs.randint(-amp, amp + 1, size=(h, w))
    elif kind == 1:
        resi = rs.laplace(0, amp / 3.0 + 0.5, size=(h, w)).astype(np.int64)
    else:
        resi = rs.randint(-amp, amp + 1, size=(h, w)); resi[rs.rand(h, w) < 0.7] = 0
    lim = (1 << bd) - 1
    shift = max(0, 15 - bd - ((int(np.log2(w)) + int(np.log2(h))) >> 1)) if int(row[11]) & 1 else 0
    return (np.clip(resi, -lim, lim) << shift).astype(np.int32)


def dqd_cases():
    """DepQuant dequantiser + inverse transform: rows trHor, trVer, w, h, bit_depth, qp, amp, seed"""
    rows = []
    rs = np.random.RandomState(606)
    seed = 17000
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (64, 64), (4, 16), (32, 8), (16, 64), (64, 32)]:
        for (th, tv) in ((0, 0), (2, 2), (1, 2)):
            if (th or tv) and (w > 32 or h > 32):
                continue
            for bd in (8, 10):
                for k in range(4):
                    rows.append([th, tv, w, h, bd, int(rs.randint(-6 * (bd - 8), 64)), int(rs.choice([1, 2, 5, 40, 600, 16000])), seed])
                    seed += 1
    return np.array(rows, dtype=np.int32)


def dqd_inputs(row, scan):
    """levels with a random last significant scan position; scan: scan position -> raster index (row pitch w)"""
    th, tv, w, h, bd, qp, amp, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    ns = min(w, 32) * min(h, 32)
    last = int(rs.randint(0, ns))
    vals = rs.randint(-amp, amp + 1, size=last + 1).astype(np.int16)
    if rs.randint(3) == 0:
        vals[rs.randint(0, 2, size=last + 1) > 0] = 0
    if vals[-1] == 0:
        vals[-1] = 1
    q = np.zeros((h, w), dtype=np.int16)
    q.reshape(-1)[scan[:last + 1]] = vals
    return q, last


def ilf_cases():
    """inverse LFNST (dequantiser, xInvLfnst, xIT on the top-left 8x8 / 4x4): rows w, h, bit_depth, qp, intra mode, lfnst index, dep quant, amp, seed"""
    rows = []
    rs = np.random.RandomState(707)
    seed = 19000
    for (w, h) in [(4, 4), (8, 8), (4, 8), (8, 4), (16, 16), (4, 16), (16, 4), (8, 16), (32, 32), (32, 8), (64, 64), (16, 64)]:
        for mode in (0, 1, 2, 10, 18, 23, 34, 35, 44, 50, 58, 66):
            for idx in (1, 2):
                bd = int(rs.choice([8, 10]))
                rows.append([w, h, bd, int(rs.randint(-6 * (bd - 8), 64)), mode, idx, int(rs.randint(0, 2)), int(rs.choice([2, 20, 300, 5000])), seed])
                seed += 1
    return np.array(rows, dtype=np.int32)


def ilf_inputs(row, scan):
    """levels as the bitstream of an LFNST TU can carry them: non-zero inside the first 8 (4x4, 8x8) or 16 scan positions only; returns levels, last position"""
    w, h, bd, qp, mode, idx, dq, amp, seed = [int(v) for v in row]
    rs = np.random.RandomState(seed)
    npos = 8 if (w, h) in ((4, 4), (8, 8)) else 16
    last = int(rs.randint(0, npos))
    vals = rs.randint(-amp, amp + 1, size=last + 1).astype(np.int16)
    if vals[-1] == 0:
        vals[-1] = -1
    q = np.zeros((h, w), dtype=np.int16)
    q.reshape(-1)[scan[:last + 1]] = vals
    return q, last
