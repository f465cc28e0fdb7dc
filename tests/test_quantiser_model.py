"""The in-register quantiser of the tensor TU engine (vvenc_b200/csrc/trquant_tc2_kernels.cuh) restates QuantCore (Quant.cpp:132-230) in a form that differs from the
reference's loop structure: per COLUMN of the coefficient block (one lane owns a column), coefficient-group granularity for the threshold pass, trimming by whole
groups, signed multiply-add levels, "last position = highest significant row + table look-up".  This file is the same algorithm in numpy, column by column, checked
against the oracle (which tests/test_oracle_vs_reference.py pins to Quant::quant): if the derivation were wrong for some shape, scan or QP, it fails here on the CPU.
It also checks the two scan properties the kernel relies on: down a column the scan position grows with the row, and the four rows of a group share a coefficient group."""
import ctypes
import math
import numpy as np
import pytest
from _libs import oracle, P

SCALES = [[26214, 23302, 20560, 18396, 16384, 14564], [18396, 16384, 14564, 13107, 11651, 10280]]      # g_quantScales, Rom.cpp


def quant_par(w, h, bd, qp, irap):
    lw = int(math.log2(w)); lh = int(math.log2(h)); sqrt2 = (lw + lh) & 1
    base = min(max(qp + 6 * (bd - 8), 0), 63 + 6 * (bd - 8)); per = base // 6; rem = base % 6
    qbits = 14 + per + 15 - bd - ((lw + lh) >> 1) - sqrt2
    scale = SCALES[sqrt2][rem]
    return scale, qbits, (171 if irap else 85) << (qbits - 9), (8 << (qbits - 1)) // (scale << 2)


def scan_tables(O, w, h):
    so = np.zeros(1024, np.int32); ns = O.orc_scan_order(w, h, P(so))
    kw, kh = min(w, 32), min(h, 32)
    inv = -np.ones((kh, kw), np.int32)
    for sp in range(ns):
        y, x = divmod(int(so[sp]), w); inv[y, x] = sp
    return inv


def device_quantiser(O, coef, w, h, bd, qp, irap):
    kw, kh = min(w, 32), min(h, 32)
    scale, qbits, add, use_thres = quant_par(w, h, bd, qp, irap)
    inv = scan_tables(O, w, h); cg = inv >> 4
    cf = coef[:kh, :kw].astype(np.int64).copy()
    amax = cg_max = init_cg = 0
    for j in range(kw):                                        # pass 1, one lane per column, then max over the lanes
        c_m = i_c = 0
        for g in range(kh // 4):
            m4 = int(np.abs(cf[4 * g:4 * g + 4, j]).max()); amax = max(amax, m4)
            if m4: i_c = int(cg[4 * g, j])                     # last assignment = maximum (group index grows down the column)
            if m4 > use_thres: c_m = int(cg[4 * g, j])
        cg_max = max(cg_max, c_m); init_cg = max(init_cg, i_c)
    trimmed = init_cg >= 1 and cg_max != init_cg
    if trimmed:
        for j in range(kw):
            for g in range(kh // 4):
                if cg[4 * g, j] > cg_max: cf[4 * g:4 * g + 4, j] = 0
    fast = qbits <= 30 and amax < 32768 and ((amax * scale + add) >> qbits) <= 32767
    q = np.zeros((h, w), np.int16); s = 0; last_q = 0
    for j in range(kw):
        hi_q = -1
        for i in range(kh):
            c = int(cf[i, j])
            if fast:
                v = (c * scale + (((1 << qbits) - 1 - add) if c < 0 else add)) >> qbits; s += abs(v)
            else:
                mag = (abs(c) * scale + add) >> qbits; s += mag; v = max(-32768, -mag) if c < 0 else min(32767, mag)
            if v: hi_q = i
            q[i, j] = v
        if hi_q >= 0: last_q = max(last_q, int(inv[hi_q, j]) + 1)
    pos = cg_max * 16 + 15
    if s == 0 and not trimmed:
        pos = 0
        for j in range(kw):
            nz = np.nonzero(cf[:, j])[0]
            if len(nz): pos = max(pos, int(inv[nz[-1], j]))
    return q, s, (last_q - 1 if s else pos)


SHAPES = [(w, h) for w in (8, 16, 32, 64) for h in (8, 16, 32, 64)]


def test_scan_properties_the_kernel_relies_on():
    O = oracle()
    for (w, h) in SHAPES:
        inv = scan_tables(O, w, h); cg = inv >> 4
        assert (inv >= 0).all(), (w, h)
        assert (np.diff(inv, axis=0) > 0).all(), (w, h)                                   # the scan position grows down every column
        assert (cg[0::4] == cg[1::4]).all() and (cg[0::4] == cg[2::4]).all() and (cg[0::4] == cg[3::4]).all(), (w, h)
        assert (np.diff(cg[::4], axis=0) >= 0).all(), (w, h)


@pytest.mark.parametrize("w,h", SHAPES)
def test_column_wise_quantiser_equals_quantcore(w, h):
    O = oracle()
    rs = np.random.RandomState(w * 100 + h)
    n_trim = n_zero = 0
    for t in range(60):
        bd = int(rs.choice([8, 10, 12])); qp = int(rs.randint(-6 * (bd - 8), 60)); irap = int(rs.randint(0, 2))
        amp = int(rs.choice([3, 20, 200, 3000, 60000]))
        coef = (rs.laplace(0, 1, size=(h, w)) * amp / (1 + np.add.outer(np.arange(h), np.arange(w))) ** rs.choice([0, 0.7, 1.5])).astype(np.int64)
        coef = np.clip(coef, -(1 << 22), 1 << 22).astype(np.int32)
        coef[:, 32:] = 0; coef[32:, :] = 0
        if rs.randint(3) == 0: coef[rs.randint(0, 2, size=(h, w)) > 0] = 0
        q = np.zeros((h, w), np.int16); s = ctypes.c_int32(); lp = ctypes.c_int32()
        assert O.orc_quant(P(coef), w, h, bd, qp, irap, P(q), ctypes.byref(s), ctypes.byref(lp)) == 0
        q2, s2, lp2 = device_quantiser(O, coef, w, h, bd, qp, irap)
        assert np.array_equal(q, q2) and s.value == s2 and lp.value == lp2, (w, h, bd, qp, irap, amp, s.value, s2, lp.value, lp2)
        n_zero += int(s.value == 0); n_trim += int(lp.value % 16 == 15)
    assert n_zero > 0 or n_trim > 0
