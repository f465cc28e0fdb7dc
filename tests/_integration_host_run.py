"""Body of tests/test_integration_host.py, run in a process of its own: the reference probe binds ONE C-ABI library per process (dlopen in
integration/RdCostB200.h), here the oracle-backed mock, while the -m gpu drop-in tests bind libvvenc_b200.so.  Prints one JSON object."""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import cases as C
import impls
from _libs import refshim, P, PO


def main(mock_path):
    R = refshim()
    R.refshim_b200_error.restype = ctypes.c_char_p
    res = {}
    assert R.refshim_install_b200(mock_path.encode()) == 0, R.refshim_b200_error()
    assert R.refshim_install_b200_search(mock_path.encode()) == 0, R.refshim_b200_error()
    dbl = ctypes.c_double
    R.refshim_pattern_search_b200.argtypes = R.refshim_pattern_search_member.argtypes
    R.refshim_row_search_b200.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int, dbl, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    fr_args = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, dbl, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    R.refshim_frac_search_member.argtypes = fr_args + [ctypes.c_int, ctypes.c_void_p]
    R.refshim_frac_search_b200.argtypes = fr_args + [ctypes.c_void_p]
    R.refshim_frac_search_b200_ex.argtypes = fr_args + [ctypes.c_int, ctypes.c_void_p]
    R.refshim_mctf_estimate_level_b200.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 4 + \
        [ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p]

    # ---- RdCost tables patched with the trampolines of RdCostB200.h (opt 2) against the AVX2 table (opt 1), golden distortion rows
    g = np.load(os.path.join(HERE, 'golden', 'golden_v1.npz'), allow_pickle=True)
    res['dist_mismatches'] = len(impls.run_dist(impls.RefImpl(opt=2), g['dist_rows'], g['dist_expect']))
    res['dist_rows'] = int(len(g['dist_rows']))

    # ---- AffineGradientSearch pointers patched by installB200( AffineGradientSearch& ) (opt 2) against the AVX2 kernels (opt 1): Sobel planes and the
    #      normal equations of every affine case row
    assert R.refshim_install_b200_affine(mock_path.encode()) == 0, R.refshim_b200_error()
    bad = 0; na = 0
    A1 = impls.RefImpl(opt=1); A2 = impls.RefImpl(opt=2)
    for row in C.affine_cases():
        w, h, ps, ds, six, seed = [int(v) for v in row]
        pred, resi, gx, gy = C.affine_inputs(row)
        for vert in (0, 1):
            bad += int(not np.array_equal(A1.sobel(vert, pred, ps, ds, w, h)[1:h - 1, 1:w - 1], A2.sobel(vert, pred, ps, ds, w, h)[1:h - 1, 1:w - 1]))
            bad += int(not np.array_equal(A1.sobel(vert, pred, ps, ds, w, h)[:, :w], A2.sobel(vert, pred, ps, ds, w, h)[:, :w]))
        e1 = A1.equal_coeff(six, resi, ps, gx, gy, ds, w, h); e2 = A2.equal_coeff(six, resi, ps, gx, gy, ds, w, h)
        bad += int(not np.array_equal(e1, e2)); na += 1
    res['affine'] = {'cases': na, 'bad': bad}

    # ---- xPatternSearchB200 / B200RowSearch against InterSearch::xPatternSearch
    sc = C.search_case()
    n = len(sc['blk']); S = sc['stride']; base = sc['margin'] * S + sc['margin']
    H = sc['org'].shape[0] - 2 * sc['margin']; W = S - 2 * sc['margin']
    res['search'] = []
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for mode in (0, 1, 2):
            for imv in (sc['imv_shift'], 2):
                a = np.zeros((n, 4), dtype=np.int32); b = np.zeros((n, 4), dtype=np.int32); c = np.zeros((n, 4), dtype=np.int32)
                args = (PO(sc['org'], base), S, PO(sc['ref'], base), S, P(sc['blk']), n, 10, mode, sc['lam'], sc['cost_scale'], imv)
                R.refshim_pattern_search_member(opt, *args, P(a))
                rc1 = R.refshim_pattern_search_b200(opt, *args, P(b))
                rc2 = R.refshim_row_search_b200(opt, PO(sc['org'], base), S, PO(sc['ref'], base), S, W, H, sc['margin'], P(sc['blk']), n, 10, mode, sc['lam'], sc['cost_scale'], imv, P(c))
                res['search'].append({'opt': opt, 'mode': mode, 'imv': imv, 'rc': [rc1, rc2], 'blocks': n, 'member_eq_b200': bool(np.array_equal(a, b)),
                                      'member_eq_rows': bool(np.array_equal(a, c)), 'err': (R.refshim_b200_error() or b'').decode() if (rc1 or rc2) else ''})

    # ---- xTZSearchB200 against InterSearch::xTZSearch: the unmodified member walking the dense SAD table; diamond / enhanced / fast settings, integer early
    #      termination, first-search stop, all sub-sampling modes; the last configuration shrinks the readable reach so that part of the walk leaves the table
    tz_args = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
               ctypes.c_int, dbl] + [ctypes.c_int] * 7 + [ctypes.c_void_p]
    R.refshim_tz_search_member.argtypes = tz_args; R.refshim_tz_search_b200.argtypes = tz_args; R.refshim_tz_search_rows_b200.argtypes = tz_args
    R.refshim_tz_search_b200_mt.argtypes = tz_args + [ctypes.c_int]
    tz = C.search_case(seed=909, W=256, H=160, margin=96)
    tS = tz['stride']; tbase = tz['margin'] * tS + tz['margin']
    rs = np.random.RandomState(5)
    tblk = []
    for (w, h) in ((8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 16), (32, 16), (4, 8), (64, 32)):
        for k in range(6):
            tblk.append((int(rs.randint(0, 256 - w + 1)), int(rs.randint(0, 160 - h + 1)), w, h, int(rs.randint(-20 * 16, 20 * 16 + 1)), int(rs.randint(-12 * 16, 12 * 16 + 1))))
    tblk = np.array(tblk, dtype=np.int32); tn = len(tblk)
    res['tz'] = []
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for (ext, fast, iet, stop, mode, rng, reach) in ((0, 0, 0, 0, 0, 24, 80), (1, 0, 0, 0, 1, 24, 80), (0, 1, 0, 1, 1, 32, 80), (0, 0, 1, 1, 2, 16, 80), (1, 1, 1, 0, 0, 24, 80),
                                                         (1, 0, 0, 0, 0, 24, 14)):
            a = np.zeros((tn, 8), dtype=np.int64); b = np.zeros((tn, 8), dtype=np.int64)
            args = (PO(tz['org'], tbase), tS, PO(tz['ref'], tbase), tS, 256, 160, reach, P(tblk), tn, 10, mode, 57.0, rng, 32, ext, fast, iet, stop, 0)
            rc1 = R.refshim_tz_search_member(opt, *args, P(a))
            rc2 = R.refshim_tz_search_b200(opt, *args, P(b))
            c = np.zeros((tn, 8), dtype=np.int64)
            rc3 = R.refshim_tz_search_rows_b200(opt, *args, P(c))                        # B200RowSearch: one launch per block size, then the same walks
            d = np.zeros((tn, 8), dtype=np.int64)
            rc4 = R.refshim_tz_search_b200_mt(opt, *args, P(d), 4)                      # four worker threads, one context / table each
            rc3 = rc3 or rc4 or int(not np.array_equal(a[:, :6], d[:, :6]))
            res['tz'].append({'opt': opt, 'cfg': [ext, fast, iet, stop, mode, rng, reach], 'rc': [rc1, rc2 or rc3],
                              'eq': bool(np.array_equal(a[:, :6], b[:, :6]) and np.array_equal(a[:, :6], c[:, :6])), 'row_hits': int(c[:, 6].sum()), 'row_misses': int(c[:, 7].sum()),
                              'hits': int(b[:, 6].sum()), 'misses': int(b[:, 7].sum()), 'moving': int((a[:, :2] != 0).any(axis=1).sum()), 'blocks': tn,
                              'err': (R.refshim_b200_error() or b'').decode() if (rc2 or rc3) else ''})

    # ---- xPatternSearchFracDIFB200 against InterSearch::xPatternSearchFracDIF (m_fastSubPel 0 and 1, SAD / SATD / fast SATD, square and rectangular PUs)
    res['frac'] = []
    rs = np.random.RandomState(23)
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        case = C.frac_case(6161 + opt)
        S = case['stride']; base = case['margin'] * S + case['margin']
        for (w, h) in ((8, 8), (16, 16), (32, 32), (64, 64), (16, 8), (8, 32), (32, 16), (8, 4), (4, 8), (4, 4), (64, 32)):
            nb = 6
            blk = np.zeros((nb, 8), dtype=np.int32)
            for k in range(nb):
                blk[k] = (int(rs.randint(0, case['W'] - w + 1)), int(rs.randint(0, case['H'] - h + 1)), w, h, int(rs.randint(-6, 7)), int(rs.randint(-6, 7)),
                          int(rs.randint(-40, 41)), int(rs.randint(-40, 41)))
            # (filter set, distortion: 0 SAD / 1 SATD / 2 fast SATD, alternative half-pel filter, m_fastSubPel)
            for (rt, had, alt, fsp) in ((2, 1, 0, 0), (0, 1, 0, 0), (1, 1, 0, 0), (2, 0, 0, 0), (0, 0, 0, 0), (2, 1, 1, 0), (2, 2, 0, 0), (2, 1, 0, 1), (2, 2, 0, 1), (2, 0, 0, 1), (2, 1, 1, 1)):
                if (w, h) == (4, 4) and rt != 2:
                    continue                                                   # 4x4 is no inter PU size; InterpolationFilter swaps in its 4x4 filter set for the longer taps there
                a = np.full((nb, 6), -7, dtype=np.int32); b = np.full((nb, 6), -7, dtype=np.int32)
                args = (PO(case['org'], base), S, PO(case['ref'], base), S, P(blk), nb, 10, 57.25, rt, had, alt)
                R.refshim_frac_search_member(opt, *args, fsp, P(a))
                rc = R.refshim_frac_search_b200_ex(opt, *args, fsp, P(b))
                res['frac'].append({'opt': opt, 'w': w, 'h': h, 'rt': rt, 'had': had, 'alt': alt, 'fsp': fsp, 'rc': rc, 'eq': bool(np.array_equal(a, b)),
                                    'err': (R.refshim_b200_error() or b'').decode() if rc else ''})
    # ---- motionEstimationLumaB200 against MCTF::motionEstimationLuma: a coarse level without predecessor, a chained level, the doubleRes final level
    assert R.refshim_install_b200_mctf(mock_path.encode()) == 0, R.refshim_b200_error()
    from test_mctf_host import _pictures
    res['mctf'] = []
    m = 128                                                                        # MCTF_PADDING
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for (W, H, seed) in ((136, 104, 31), (96, 64, 32)):
            org, ref, S = _pictures(seed + opt, W, H, m, shift=(2, -3))
            base = m * S + m

            def level(use, bs, prev, factor, dbl, low, pattern, unit=16):
                oxn, oyn = (W + bs - 8) // bs, (H + bs - 8) // bs
                out = np.zeros((oyn, oxn, 4), dtype=np.int32); ov = np.zeros((oyn, oxn), dtype=np.float64)
                pv = None if prev is None else np.ascontiguousarray(prev[:, :, :2])
                rc = R.refshim_mctf_estimate_level_b200(opt, use, PO(org, base), S, PO(ref, base), S, W, H, 10, bs, None if pv is None else P(pv), 0 if pv is None else pv.shape[1],
                                                        0 if pv is None else pv.shape[0], factor, int(dbl), low, unit, pattern, P(out), P(ov))
                return rc, out, ov

            for pattern in (0, 1, 2):
                for low in (0, 1):
                    rc0, a0, _ = level(0, 32, None, 2, False, low, pattern)
                    rc1, b0, _ = level(1, 32, None, 2, False, low, pattern)
                    rc2, a1, _ = level(0, 16, a0, 1, False, low, pattern)
                    rc3, b1, _ = level(1, 16, a0, 1, False, low, pattern)
                    rc4, a2, oa = level(0, 16, a0, 1, True, low, pattern)
                    rc5, b2, ob = level(1, 16, a0, 1, True, low, pattern)
                    res['mctf'].append({'opt': opt, 'W': W, 'H': H, 'pattern': pattern, 'low': low, 'rc': [rc0, rc1, rc2, rc3, rc4, rc5],
                                        'eq': [bool(np.array_equal(a0, b0)), bool(np.array_equal(a1, b1)), bool(np.array_equal(a2, b2)), bool(np.array_equal(oa, ob))],
                                        'moving': int((a2[:, :, :2] != 0).any(axis=2).sum()), 'blocks': int(a2.shape[0] * a2.shape[1]),
                                        'err': (R.refshim_b200_error() or b'').decode() if any([rc1, rc3, rc5]) else ''})
    # ---- bilateralFilterB200 against MCTF::bilateralFilter (whole small luma pictures: unit 8 / 16, QP on both sides of the planar-correction threshold,
    #      6- and 4-tap apply filters, 2..8 neighbour pictures, both strength rows, clipped edge blocks)
    res['mctf_apply'] = []
    rs = np.random.RandomState(404)
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for (W, H, unit, nrefs, qp, tap4, reorder) in ((96, 64, 16, 4, 22, 0, 1), (72, 40, 8, 6, 40, 0, 1), (64, 48, 16, 2, 32, 1, 0), (48, 32, 8, 8, 27, 0, 1), (80, 56, 32, 3, 30, 0, 0)):
            base_ = rs.randint(0, 1024, size=(H + 8, W + 8))
            sm = (base_ + np.roll(base_, 1, 0) + np.roll(base_, 1, 1) + np.roll(base_, (1, 1), (0, 1))) // 4
            org = np.ascontiguousarray(sm[4:4 + H, 4:4 + W].astype(np.int16))
            refs = []
            for a in ([3, 12, 60, 300] * 2)[:nrefs]:
                dy = int(rs.randint(-1, 2)); dx = int(rs.randint(-1, 2))
                refs.append(np.ascontiguousarray(np.clip(sm[4 + dy:4 + dy + H, 4 + dx:4 + dx + W] + rs.randint(-a, a + 1, size=org.shape), 0, 1023).astype(np.int16)))
            wb, hb = (W + unit - 1) // unit, (H + unit - 1) // unit
            mv = np.zeros((nrefs, hb * wb, 4), dtype=np.int32)
            mv[..., 0] = rs.randint(-80, 81, size=(nrefs, hb * wb)); mv[..., 1] = rs.randint(-80, 81, size=(nrefs, hb * wb))
            mv[..., 2] = rs.choice([3, 20, 49, 50, 75, 100, 101, 400], size=(nrefs, hb * wb)); mv[..., 3] = rs.choice([0, 1, 5, 22, 60], size=(nrefs, hb * wb))
            idx = np.array([i % 6 for i in range(nrefs)], dtype=np.int32)
            ptrs = (ctypes.c_void_p * nrefs)(*[r_.ctypes.data for r_ in refs])
            a_ = np.zeros((H, W), dtype=np.int16); b_ = np.zeros((H, W), dtype=np.int16)
            R.refshim_mctf_bilateral_filter(opt, P(org), ptrs, nrefs, P(mv), P(idx), W, H, 10, unit, qp, ctypes.c_double(0.95), reorder, tap4, P(a_), None, None)
            rc = R.refshim_mctf_bilateral_filter_b200(opt, P(org), ptrs, nrefs, P(mv), P(idx), W, H, 10, unit, qp, ctypes.c_double(0.95), reorder, tap4, P(b_))
            res['mctf_apply'].append({'opt': opt, 'W': W, 'H': H, 'unit': unit, 'refs': nrefs, 'qp': qp, 'rc': rc, 'eq': bool(np.array_equal(a_, b_)),
                                      'changed': bool(np.any(a_ != org)), 'err': (R.refshim_b200_error() or b'').decode() if rc else ''})

    # ---- the same on 4:2:0 pictures: all three components (chroma: half-size units, vectors scaled by the sub-sampling, its own weight and sigma)
    R.refshim_mctf_bilateral_filter420.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 5 + \
        [ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    res['mctf_apply420'] = []
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for (W, H, unit, nrefs, qp, tap4) in ((96, 64, 16, 4, 22, 0), (128, 96, 16, 3, 40, 1), (64, 64, 32, 2, 30, 0), (80, 48, 16, 8, 32, 0)):
            def i420(seed_shift, noise):
                planes = []
                for (pw, ph) in ((W, H), (W // 2, H // 2), (W // 2, H // 2)):
                    b0 = rs.randint(0, 1024, size=(ph + 8, pw + 8))
                    sm_ = (b0 + np.roll(b0, 1, 0) + np.roll(b0, 1, 1) + np.roll(b0, (1, 1), (0, 1))) // 4
                    planes.append(np.clip(sm_[4:4 + ph, 4:4 + pw] + rs.randint(-noise, noise + 1, size=(ph, pw)), 0, 1023).astype(np.int16).reshape(-1))
                return np.ascontiguousarray(np.concatenate(planes))
            org = i420(0, 0)
            refs = [np.ascontiguousarray(np.clip(org.astype(np.int32) + rs.randint(-a, a + 1, size=org.shape), 0, 1023).astype(np.int16)) for a in ([3, 12, 60, 300] * 2)[:nrefs]]
            wb, hb = (W + unit - 1) // unit, (H + unit - 1) // unit
            mv = np.zeros((nrefs, hb * wb, 4), dtype=np.int32)
            mv[..., 0] = rs.randint(-80, 81, size=(nrefs, hb * wb)); mv[..., 1] = rs.randint(-80, 81, size=(nrefs, hb * wb))
            mv[..., 2] = rs.choice([3, 20, 49, 50, 75, 100, 101, 400], size=(nrefs, hb * wb)); mv[..., 3] = rs.choice([0, 1, 5, 22, 60], size=(nrefs, hb * wb))
            idx = np.array([i % 6 for i in range(nrefs)], dtype=np.int32)
            ptrs = (ctypes.c_void_p * nrefs)(*[r_.ctypes.data for r_ in refs])
            a_ = np.zeros_like(org); b_ = np.zeros_like(org)
            rc1 = R.refshim_mctf_bilateral_filter420(opt, 0, P(org), ptrs, nrefs, P(mv), P(idx), W, H, 10, unit, qp, 0.95, 1, tap4, P(a_))
            rc2 = R.refshim_mctf_bilateral_filter420(opt, 1, P(org), ptrs, nrefs, P(mv), P(idx), W, H, 10, unit, qp, 0.95, 1, tap4, P(b_))
            ny = W * H
            res['mctf_apply420'].append({'opt': opt, 'W': W, 'H': H, 'unit': unit, 'refs': nrefs, 'qp': qp, 'rc': [rc1, rc2], 'eq_luma': bool(np.array_equal(a_[:ny], b_[:ny])),
                                         'eq_chroma': bool(np.array_equal(a_[ny:], b_[ny:])), 'chroma_changed': bool(np.any(a_[ny:] != org[ny:])),
                                         'err': (R.refshim_b200_error() or b'').decode() if (rc1 or rc2) else ''})

    # ---- xTQuantB200 / invTransformNxNB200 against TrQuant::xT + Quant::quant / Quant::dequant + xIT on the probe's TransformUnit rig: every case row of the
    #      parity tables (all shapes, DCT-II / DST-VII / DCT-VIII pairs, 8 and 10 bit, strided residuals, both slice types, xNeedRDOQ with and without depQuant)
    assert R.refshim_install_b200_tu(mock_path.encode()) == 0, R.refshim_b200_error()
    I32 = ctypes.c_int32
    bad = []; nfwd = 0
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for row in C.tq_cases()[opt::2]:
            th, tv, w, h, st, amp, qp, irap, bd, seed = [int(v) for v in row]
            resi = C.tq_inputs(row)
            dq = seed & 1
            ca = np.zeros((h, w), dtype=np.int32); qa = np.zeros((h, w), dtype=np.int16); sa = I32(); la = I32()
            cb = np.zeros((h, w), dtype=np.int32); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); lb = I32(); nb = I32()
            assert R.refshim_transform_quant(th, tv, P(resi), st, w, h, bd, qp, irap, P(ca), P(qa), ctypes.byref(sa), ctypes.byref(la)) == 0
            na = R.refshim_need_rdoq(P(ca), w, h, bd, qp, dq)
            rc = R.refshim_transform_quant_b200(th, tv, P(resi), st, w, h, bd, qp, irap, dq, P(cb), P(qb), ctypes.byref(sb), ctypes.byref(lb), ctypes.byref(nb))
            nfwd += 1
            if rc or not (np.array_equal(ca, cb) and np.array_equal(qa, qb) and sa.value == sb.value and la.value == lb.value and int(na) == nb.value):
                bad.append(['fwd', opt] + [int(v) for v in row] + [rc])
    res['tu_fwd'] = {'cases': nfwd, 'bad': bad[:5]}
    # the same pair with slice->signDataHidingEnabled: Quant::quant ends in xSignBitHidingHDQ, the library hides on the device (vvb_tu_par.sign_hiding)
    bad = []; nsdh = 0; hid = 0
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for row in C.tq_cases()[1 - opt::2]:
            th, tv, w, h, st, amp, qp, irap, bd, seed = [int(v) for v in row]
            resi = C.tq_inputs(row)
            ca = np.zeros((h, w), dtype=np.int32); qa = np.zeros((h, w), dtype=np.int16); sa = I32(); la = I32()
            cb = np.zeros((h, w), dtype=np.int32); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); lb = I32(); nb = I32()
            q0 = np.zeros((h, w), dtype=np.int16)
            assert R.refshim_transform_quant(th, tv, P(resi), st, w, h, bd, qp, irap, P(ca), P(q0), ctypes.byref(sa), ctypes.byref(la)) == 0
            assert R.refshim_transform_quant_sdh(th, tv, P(resi), st, w, h, bd, qp, irap, 1, P(ca), P(qa), ctypes.byref(sa), ctypes.byref(la)) == 0
            rc = R.refshim_transform_quant_b200_sdh(th, tv, P(resi), st, w, h, bd, qp, irap, 0, 1, P(cb), P(qb), ctypes.byref(sb), ctypes.byref(lb), ctypes.byref(nb))
            nsdh += 1; hid += int(not np.array_equal(q0, qa))
            if rc or not (np.array_equal(ca, cb) and np.array_equal(qa, qb) and sa.value == sb.value and la.value == lb.value):
                bad.append(['sdh', opt] + [int(v) for v in row] + [rc])
    res['tu_fwd_sdh'] = {'cases': nsdh, 'levels_changed_by_hiding': hid, 'bad': bad[:5]}
    # intra TUs with an LFNST index: xT (LFNST zero-out) + xFwdLfnst + Quant::quant against xTQuantB200 with vvb_tu_par.lfnst_*
    bad = []; nlf = 0
    rs2 = np.random.RandomState(77)
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for (w, h) in ((4, 4), (8, 8), (4, 8), (16, 4), (16, 16), (8, 32), (32, 32), (64, 64), (64, 16)):
            for mode in (0, 1, 2, 18, 34, 35, 50, 66):
                for idx in (1, 2):
                    amp = int(rs2.choice([1023, 200, 30])); qp = int(rs2.randint(18, 46)); irap = int(rs2.randint(0, 2)); sh = int(rs2.randint(0, 2))
                    resi = rs2.randint(-amp, amp + 1, size=(h, w)).astype(np.int16)
                    ca = np.zeros((h, w), dtype=np.int32); qa = np.zeros((h, w), dtype=np.int16); sa = I32(); la = I32(); na = I32(); st2 = np.zeros(2, dtype=np.int32)
                    cb = np.zeros((h, w), dtype=np.int32); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); lb = I32(); nb = I32()
                    assert R.refshim_transform_quant_lfnst(P(resi), w, w, h, 10, qp, irap, sh, mode, idx, P(ca), P(qa), ctypes.byref(sa), ctypes.byref(la), ctypes.byref(na), P(st2)) == 0
                    rc = R.refshim_transform_quant_lfnst_b200(P(resi), w, w, h, 10, qp, irap, sh, mode, idx, P(cb), P(qb), ctypes.byref(sb), ctypes.byref(lb), ctypes.byref(nb))
                    nlf += 1
                    if rc or not (np.array_equal(ca, cb) and np.array_equal(qa, qb) and sa.value == sb.value and la.value == lb.value and na.value == nb.value):
                        bad.append(['lfnst', opt, w, h, mode, idx, qp, irap, sh, rc])
    res['tu_fwd_lfnst'] = {'cases': nlf, 'bad': bad[:5]}
    bad = []; ninv = 0
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for row in C.itq_cases()[opt::2]:
            th, tv, w, h, st, kind, qp, bd, seed = [int(v) for v in row]
            q = C.itq_inputs(row)
            ra = np.zeros((h, st), dtype=np.int16); rb = np.zeros((h, st), dtype=np.int16)
            assert R.refshim_inv_transform_quant(th, tv, P(q), w, h, bd, qp, None, P(ra), st) == 0
            rc = R.refshim_inv_transform_quant_b200(th, tv, P(q), w, h, bd, qp, P(rb), st)
            ninv += 1
            if rc or not np.array_equal(ra, rb):
                bad.append(['inv', opt] + [int(v) for v in row] + [rc])
    res['tu_inv'] = {'cases': ninv, 'bad': bad[:5]}
    # transform skip and chroma components through xTQuantB200 / invTransformNxNB200 (vvb_tu_par.transform_skip, input_bit_depth_delta, is_chroma)
    bad = []; nts = 0; ninv = 0
    R.refshim_set_simd(b'AVX2')
    for row in C.ts_cases():
        w, h, st, bd, amp, qp, irap, sh, dq, ts, delta, comp, seed = [int(v) for v in row]
        resi = C.ts_inputs(row)
        ca = np.zeros((h, w), dtype=np.int32); qa = np.zeros((h, w), dtype=np.int16); sa = I32(); la = I32(); na = I32()
        cb = np.zeros((h, w), dtype=np.int32); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); lb = I32(); nb = I32()
        assert R.refshim_transform_quant_ts(P(resi), st, w, h, bd, qp, irap, sh, dq, ts, delta, comp, P(ca), P(qa), ctypes.byref(sa), ctypes.byref(la), ctypes.byref(na)) == 0
        rc = R.refshim_transform_quant_ts_b200(P(resi), st, w, h, bd, qp, irap, sh, dq, ts, delta, comp, P(cb), P(qb), ctypes.byref(sb), ctypes.byref(lb), ctypes.byref(nb))
        nts += 1
        if rc or not (np.array_equal(ca, cb) and np.array_equal(qa, qb) and sa.value == sb.value and la.value == lb.value and na.value == nb.value):
            bad.append(['ts'] + [int(v) for v in row] + [rc, (R.refshim_b200_error() or b'').decode() if rc else ''])
        if ts and sa.value > 0 and comp == 0:
            ra = np.zeros((h, st), dtype=np.int16); rb = np.zeros((h, st), dtype=np.int16)
            assert R.refshim_inv_transform_quant_ts(P(qa), w, h, bd, qp, delta, None, P(ra), st) == 0
            rc = R.refshim_inv_transform_quant_ts_b200(P(qa), w, h, bd, qp, delta, P(rb), st)
            ninv += 1
            if rc or not np.array_equal(ra[:, :w], rb[:, :w]):
                bad.append(['ts_inv'] + [int(v) for v in row] + [rc])
    res['tu_ts_chroma'] = {'cases': nts, 'inverse_cases': ninv, 'bad': bad[:5]}
    # DepQuant::dequant + xIT against invTransformNxNB200 with par.dep_quant
    bad = []; ndd = 0
    from _libs import oracle as _orc
    for row in C.dqd_cases():
        th, tv, w, h, bd, qp, amp, seed = [int(v) for v in row]
        so = np.zeros(1024, np.int32); _orc().orc_scan_order(w, h, P(so))
        q, last = C.dqd_inputs(row, so)
        ra = np.zeros((h, w), dtype=np.int16); rb = np.zeros((h, w), dtype=np.int16)
        assert R.refshim_inv_transform_quant_dq(th, tv, P(q), last, w, h, bd, qp, None, P(ra), w) == 0
        rc = R.refshim_inv_transform_quant_dq_b200(th, tv, P(q), last, w, h, bd, qp, P(rb), w)
        ndd += 1
        if rc or not np.array_equal(ra, rb):
            bad.append(['dqd'] + [int(v) for v in row] + [rc])
    res['tu_inv_dq'] = {'cases': ndd, 'bad': bad[:5]}
    # invTransformNxN of LFNST TUs (dequantiser, xInvLfnst, xIT) against invTransformNxNB200 with vvb_tu_par.lfnst_*
    bad = []; nil = 0
    for opt in (0, 1):
        R.refshim_set_simd(b'AVX2' if opt else b'SCALAR')
        for row in C.ilf_cases()[opt::2]:
            w, h, bd, qp, mode, idx, dq, amp, seed = [int(v) for v in row]
            so = np.zeros(1024, np.int32); _orc().orc_scan_order(w, h, P(so))
            q, last = C.ilf_inputs(row, so)
            ra = np.zeros((h, w), dtype=np.int16); rb = np.zeros((h, w), dtype=np.int16); st2 = np.zeros(2, dtype=np.int32)
            assert R.refshim_inv_transform_quant_lfnst(P(q), w, h, bd, qp, dq, last, mode, idx, None, P(ra), w, P(st2)) == 0
            rc = R.refshim_inv_transform_quant_lfnst_b200(P(q), w, h, bd, qp, dq, last, mode, idx, P(rb), w)
            nil += 1
            if rc or not np.array_equal(ra, rb):
                bad.append(['ilf'] + [int(v) for v in row] + [rc, (R.refshim_b200_error() or b'').decode() if rc else ''])
    res['tu_inv_lfnst'] = {'cases': nil, 'bad': bad[:5]}
    # DepQuant::xQuantDQ against xQuantDQB200 (rate tables from the rig's CABAC contexts through the public RateEstimator accessors, trellis in the bound library)
    bad = []; ndq = 0; nz = 0
    R.refshim_set_simd(b'AVX2')
    for row in C.dq_cases():
        w, h, bd, qp, lam1000, scale, decay10, mts, lf, sbt, intra, init_id, seed = [int(v) for v in row]
        coef = C.dq_inputs(row)
        qa = np.zeros((h, w), dtype=np.int16); sa = I32(); la = I32(); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); lb = I32()
        assert R.refshim_dep_quant(P(coef), w, h, bd, qp, mts, intra, lf, sbt, lam1000 / 1000.0, 8, 1, qp, init_id, P(qa), ctypes.byref(sa), ctypes.byref(la), None, None) == 0
        rc = R.refshim_dep_quant_b200(P(coef), w, h, bd, qp, mts, intra, lf, sbt, lam1000 / 1000.0, 8, qp, init_id, P(qb), ctypes.byref(sb), ctypes.byref(lb))
        ndq += 1; nz += int(la.value >= 0)
        if rc or not (np.array_equal(qa, qb) and sa.value == sb.value and la.value == lb.value):
            bad.append(['dq'] + [int(v) for v in row] + [rc, (R.refshim_b200_error() or b'').decode() if rc else ''])
    res['dep_quant'] = {'cases': ndq, 'non_empty': nz, 'bad': bad[:5]}
    bad = []; ndc = 0
    for row in C.dq_chroma_cases():
        w, h, bd, qp, lam1000, scale, decay10, lf, intra, init_id, seed = [int(v) for v in row]
        coef = C.dq_chroma_inputs(row)
        qa = np.zeros((h, w), dtype=np.int16); sa = I32(); la = I32(); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); lb = I32()
        assert R.refshim_dep_quant_comp(1, P(coef), w, h, bd, qp, 0, intra, lf, 0, lam1000 / 1000.0, 8, 1, qp, init_id, P(qa), ctypes.byref(sa), ctypes.byref(la), None, None) == 0
        rc = R.refshim_dep_quant_b200_comp(1, P(coef), w, h, bd, qp, 0, intra, lf, 0, lam1000 / 1000.0, 8, qp, init_id, P(qb), ctypes.byref(sb), ctypes.byref(lb))
        ndc += 1
        if rc or not (np.array_equal(qa, qb) and sa.value == sb.value and la.value == lb.value):
            bad.append(['dq_chroma'] + [int(v) for v in row] + [rc, (R.refshim_b200_error() or b'').decode() if rc else ''])
    res['dep_quant_chroma'] = {'cases': ndc, 'bad': bad[:5]}
    print('RESULT ' + json.dumps(res))


def main_rdoq(lib_path):
    """QuantRDOQ2::xRateDistOptQuant (member, AVX2 build) against xRateDistOptQuantB200 (fractional bits from the rig's CABAC contexts, the member's own last-position
    table incl. the Cr-after-coded-Cb reuse, level decisions in the bound library): every row of cases.rdoq_cases().  A run of its own (argument `rdoq`) so that the
    long-standing binding run keeps its exact shape."""
    R = refshim()
    R.refshim_b200_error.restype = ctypes.c_char_p
    assert R.refshim_install_b200(lib_path.encode()) == 0, R.refshim_b200_error()
    assert R.refshim_install_b200_tu(lib_path.encode()) == 0, R.refshim_b200_error()
    R.refshim_set_simd(b'AVX2')
    I32 = ctypes.c_int32
    bad = []; n = 0; nz = 0; nsh = 0
    for row in C.rdoq_cases():
        w, h, bd, qp, lam1000, scale, decay10, comp, lf, sbt, intra, sh, cb, thr, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_inputs(row)
        qa = np.zeros((h, w), dtype=np.int16); sa = I32(); la = I32(); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); lb = I32()
        assert R.refshim_rdoq(comp, P(coef), w, h, bd, qp, intra, lf, sbt, sh, cb, lam1000 / 1000.0, thr, qp, init_id, P(qa), ctypes.byref(sa), ctypes.byref(la), None, None) == 0
        rc = R.refshim_rdoq_b200(comp, P(coef), w, h, bd, qp, intra, lf, sbt, sh, cb, lam1000 / 1000.0, thr, qp, init_id, P(qb), ctypes.byref(sb), ctypes.byref(lb))
        n += 1; nz += int(la.value >= 0); nsh += int(sh and la.value >= 0)
        if rc or not (np.array_equal(qa, qb) and sa.value == sb.value and la.value == lb.value):
            bad.append(['rdoq'] + [int(v) for v in row] + [rc, (R.refshim_b200_error() or b'').decode() if rc else ''])
    # QuantRDOQ::rateDistOptQuantTS against rateDistOptQuantTSB200: every row of cases.rdoq_ts_cases()
    tbad = []; tn = 0; tnz = 0
    for row in C.rdoq_ts_cases():
        w, h, bd, qp, lam1000, amp, kind, comp, intra, delta, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_ts_inputs(row)
        qa = np.zeros((h, w), dtype=np.int16); sa = I32(); qb = np.zeros((h, w), dtype=np.int16); sb = I32()
        cq = qp if qp > 16 else 27
        assert R.refshim_rdoq_ts(comp, P(coef), w, h, bd, qp, delta, intra, lam1000 / 1000.0, cq, init_id, P(qa), ctypes.byref(sa), None, None, None) == 0
        rc = R.refshim_rdoq_ts_b200(comp, P(coef), w, h, bd, qp, delta, intra, lam1000 / 1000.0, cq, init_id, P(qb), ctypes.byref(sb))
        tn += 1; tnz += int(sa.value > 0)
        if rc or not (np.array_equal(qa, qb) and sa.value == sb.value):
            tbad.append(['rdoq_ts'] + [int(v) for v in row] + [rc, (R.refshim_b200_error() or b'').decode() if rc else ''])
    # QuantRDOQ::forwardRDPCM against forwardRDPCMB200, and the inverse path of the BDPCM levels (Quant::dequant incl. invResDPCM + xITransformSkip) against invTransformNxNB200
    bbad = []; bn = 0; bnz = 0
    for row in C.rdoq_ts_cases():
        w, h, bd, qp, lam1000, amp, kind, comp, intra, delta, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_ts_inputs(row); dm = 1 + (seed & 1)
        qa = np.zeros((h, w), dtype=np.int16); sa = I32(); qb = np.zeros((h, w), dtype=np.int16); sb = I32(); ra = np.zeros((h, w), dtype=np.int16); rb = np.zeros((h, w), dtype=np.int16)
        cq = qp if qp > 16 else 27
        assert R.refshim_rdoq_bdpcm(comp, P(coef), w, h, bd, qp, delta, 1, dm, lam1000 / 1000.0, cq, init_id, P(qa), ctypes.byref(sa), None) == 0
        rc = R.refshim_rdoq_bdpcm_b200(comp, P(coef), w, h, bd, qp, delta, dm, lam1000 / 1000.0, cq, init_id, P(qb), ctypes.byref(sb), P(ra), P(rb))
        bn += 1; bnz += int(sa.value > 0)
        if rc or not (np.array_equal(qa, qb) and sa.value == sb.value and np.array_equal(ra, rb)):
            bbad.append(['rdoq_bdpcm'] + [int(v) for v in row] + [rc, (R.refshim_b200_error() or b'').decode() if rc else ''])
    print('RESULT ' + json.dumps({'rdoq': {'cases': n, 'non_empty': nz, 'non_empty_with_hiding': nsh, 'bad': bad[:5]}, 'rdoq_ts': {'cases': tn, 'non_empty': tnz, 'bad': tbad[:5]},
                                  'rdoq_bdpcm': {'cases': bn, 'non_empty': bnz, 'bad': bbad[:5]}}))


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[2] == 'rdoq':
        main_rdoq(sys.argv[1])
    else:
        main(sys.argv[1])
