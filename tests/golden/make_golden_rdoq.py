#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v6_rdoq.npz from the UNMODIFIED reference (oracle/_ref): for every row of cases.rdoq_cases() the fractional bits the reference
read from its CABAC contexts (vvb_rdoq_rates layout), the per-call constants, and the levels / absSum / lastPos of QuantRDOQ2::xRateDistOptQuant; the scalar and the
SIMD build of the routine (its threshold pre-test has an SSE form, QuantRDOQ2.cpp:601-637) are required to agree at generation time.  Also every row of
cases.rdoq_ts_cases(): QuantRDOQ::rateDistOptQuantTS (transform-skipped TUs) with the fractional bits of the transform-skip context sets.
Run in the build container only:  python tests/golden/make_golden_rdoq.py"""
import ctypes, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as C
from _libs import refshim, P


def main():
    R = refshim()
    rows = C.rdoq_cases()
    out = {'cases': rows}
    rates = np.zeros((len(rows), 190), dtype=np.int32); consts = np.zeros((len(rows), 7), dtype=np.int32); meta = np.zeros((len(rows), 2), dtype=np.int32)
    for i, row in enumerate(rows):
        w, h, bd, qp, lam1000, scale, decay10, comp, lf, sbt, intra, sh, cb, thr, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_inputs(row)
        lv = []
        for simd in (b'SCALAR', b'AVX2'):
            R.refshim_set_simd(simd)
            q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
            assert R.refshim_rdoq(comp, P(coef), w, h, bd, qp, intra, lf, sbt, sh, cb, lam1000 / 1000.0, thr, qp, init_id, P(q), ctypes.byref(s), ctypes.byref(l), P(rates[i]), P(consts[i])) == 0
            lv.append((q, s.value, l.value))
        assert np.array_equal(lv[0][0], lv[1][0]) and lv[0][1:] == lv[1][1:], i
        meta[i] = lv[1][1:]; out['q_%d' % i] = lv[1][0]
    out['rates'] = rates; out['consts'] = consts; out['meta'] = meta
    # transform-skipped TUs: QuantRDOQ::rateDistOptQuantTS
    R.refshim_set_simd(b'AVX2')
    trow = C.rdoq_ts_cases()
    trates = np.zeros((len(trow), 44), dtype=np.int32); tconsts = np.zeros((len(trow), 3), dtype=np.int32); terr = np.zeros(len(trow), dtype=np.float64); tsum = np.zeros(len(trow), dtype=np.int32)
    for i, row in enumerate(trow):
        w, h, bd, qp, lam1000, amp, kind, comp, intra, delta, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_ts_inputs(row)
        q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); e = ctypes.c_double()
        assert R.refshim_rdoq_ts(comp, P(coef), w, h, bd, qp, delta, intra, lam1000 / 1000.0, qp if qp > 16 else 27, init_id, P(q), ctypes.byref(s), P(trates[i]), P(tconsts[i]), ctypes.byref(e)) == 0
        tsum[i] = s.value; terr[i] = e.value; out['tsq_%d' % i] = q
    # BDPCM TUs: QuantRDOQ::forwardRDPCM on the same rows, direction 1 + (seed & 1), intra CU
    bsum = np.zeros(len(trow), dtype=np.int32)
    for i, row in enumerate(trow):
        w, h, bd, qp, lam1000, amp, kind, comp, intra, delta, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_ts_inputs(row)
        q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); rt = np.zeros(44, dtype=np.int32)
        assert R.refshim_rdoq_bdpcm(comp, P(coef), w, h, bd, qp, delta, 1, 1 + (seed & 1), lam1000 / 1000.0, qp if qp > 16 else 27, init_id, P(q), ctypes.byref(s), P(rt)) == 0
        assert np.array_equal(rt, trates[i])
        bsum[i] = s.value; out['bdq_%d' % i] = q
    out['bd_abs_sum'] = bsum
    out['ts_cases'] = trow; out['ts_rates'] = trates; out['ts_consts'] = tconsts; out['ts_err_scale'] = terr; out['ts_abs_sum'] = tsum
    path = os.path.join(HERE, 'golden_v6_rdoq.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(rows), 'cases,', int((meta[:, 1] >= 0).sum()), 'non-empty;', len(trow), 'transform-skip cases,', int((tsum > 0).sum()), 'non-empty,', int((bsum > 0).sum()), 'non-empty with BDPCM,', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
