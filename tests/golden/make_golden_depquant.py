#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v5_depquant.npz from the UNMODIFIED reference (oracle/_ref): for every row of cases.dq_cases() the RateEstimator tables the
reference derived from its CABAC contexts, the Quantizer constants, and the levels / absSum / lastPos of DepQuant::xQuantDQ with the scalar and with the x86 members
(they differ only for levels above 127, see include/vvenc_b200.h).  Run in the build container only:  python tests/golden/make_golden_depquant.py"""
import ctypes, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as C
from _libs import refshim, P


def main():
    R = refshim()
    R.refshim_dep_quant.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_double] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 5
    rows = C.dq_cases()
    out = {'cases': rows}
    rates = np.zeros((len(rows), 266), dtype=np.int32); consts = np.zeros((len(rows), 9), dtype=np.int64)
    meta = np.zeros((len(rows), 2, 2), dtype=np.int32)       # [case][member set][abs_sum, last_pos]
    differ = 0
    for i, row in enumerate(rows):
        w, h, bd, qp, lam1000, scale, decay10, mts, lf, sbt, intra, init_id, seed = [int(v) for v in row]
        coef = C.dq_inputs(row)
        lv = []
        for opt in (0, 1):
            q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
            assert R.refshim_dep_quant(P(coef), w, h, bd, qp, mts, intra, lf, sbt, lam1000 / 1000.0, 8, opt, qp, init_id, P(q), ctypes.byref(s), ctypes.byref(l),
                                       P(rates[i]), P(consts[i])) == 0
            meta[i, opt] = (s.value, l.value); lv.append(q)
        out['q_scalar_%d' % i] = lv[0]
        if not np.array_equal(lv[0], lv[1]):
            out['q_x86_%d' % i] = lv[1]; differ += 1
    out['rates'] = rates; out['consts'] = consts; out['meta'] = meta
    # chroma components (Cb of a 4:4:4 rig): the chroma context sets and context offsets
    R.refshim_dep_quant_comp.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_double] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 5
    crow = C.dq_chroma_cases()
    crates = np.zeros((len(crow), 266), dtype=np.int32); cmeta = np.zeros((len(crow), 2), dtype=np.int32)
    for i, row in enumerate(crow):
        w, h, bd, qp, lam1000, scale, decay10, lf, intra, init_id, seed = [int(v) for v in row]
        coef = C.dq_chroma_inputs(row)
        lv = []
        for opt in (0, 1):
            q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
            assert R.refshim_dep_quant_comp(1, P(coef), w, h, bd, qp, 0, intra, lf, 0, lam1000 / 1000.0, 8, opt, qp, init_id, P(q), ctypes.byref(s), ctypes.byref(l), P(crates[i]), None) == 0
            lv.append(q)
        cmeta[i] = (s.value, l.value); out['cq_%d' % i] = lv[1]            # the x86 members (what the library follows by default)
        if not np.array_equal(lv[0], lv[1]):
            out['cq_scalar_%d' % i] = lv[0]
    out['chroma_cases'] = crow; out['chroma_rates'] = crates; out['chroma_meta'] = cmeta
    path = os.path.join(HERE, 'golden_v5_depquant.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, len(rows), 'cases,', differ, 'with scalar != x86 members,', int((meta[:, 0, 1] >= 0).sum()), 'non-empty,', os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
