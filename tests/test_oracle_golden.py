"""CPU-only: the oracle (oracle/oracle.c) against the committed golden vectors, which were produced by the
unmodified reference (scalar == AVX2 asserted at generation time, tests/golden/make_golden.py)."""
import numpy as np
import cases as C
import impls


def test_case_tables_match_fixture(golden):
    # the fixture is only meaningful if the seeded case generators still produce the rows it was built from
    assert np.array_equal(golden['dist_rows'], C.dist_cases())
    assert np.array_equal(golden['tq_rows'], C.tq_cases())
    assert np.array_equal(golden['mctf_rows'], C.mctf_cases())
    assert np.array_equal(golden['aff_rows'], C.affine_cases())


def test_oracle_dist_golden(golden):
    assert impls.run_dist(impls.OracleImpl(), golden['dist_rows'], golden['dist_expect']) == []


def test_oracle_transform_quant_golden(golden):
    assert impls.run_tq(impls.OracleImpl(), golden['tq_rows'], golden['tq_coef'], golden['tq_q'], golden['tq_meta']) == []


def test_oracle_mctf_golden(golden):
    assert impls.run_mctf(impls.OracleImpl(), golden['mctf_rows'], golden['mctf_expect']) == []


def test_oracle_affine_golden(golden):
    assert impls.run_affine(impls.OracleImpl(), golden['aff_rows'], golden['aff_sobel'], golden['aff_eq']) == []


def test_oracle_full_search_golden(golden):
    sc = C.search_case()
    O = impls.OracleImpl()
    for ss in (0, 1):
        assert np.array_equal(O.full_search(sc, ss), golden['search_best_ss%d' % ss])


def test_oracle_mv_rate_golden(golden):
    O = impls.OracleImpl()
    for i, row in enumerate(golden['mv_rows']):
        a = [int(v) for v in row]
        assert O.mv_bits(*a) == int(golden['mv_bits'][i])
        assert O.mv_cost(57.25 + i, *a) == int(golden['mv_cost'][i])


def test_tu_case_tables_match_fixture(golden_tu):
    assert np.array_equal(golden_tu['itq_rows'], C.itq_cases())
    assert np.array_equal(golden_tu['rt_rows'], C.rt_cases())


def test_oracle_inverse_path_golden(golden_tu):
    assert impls.run_itq(impls.OracleImpl(), golden_tu['itq_rows'], golden_tu['itq_coef'], golden_tu['itq_resi']) == []


def test_oracle_tu_roundtrip_golden(golden_tu):
    assert impls.run_rt(impls.OracleImpl(), golden_tu['rt_rows'], golden_tu['rt_q'], golden_tu['rt_reco'], golden_tu['rt_meta']) == []


def test_oracle_mctf_apply_golden(golden_mctf_apply):
    import ctypes
    from _libs import oracle, P
    O = oracle(); O.orc_mctf_calc_var.restype = ctypes.c_double
    for k, (seed, W, H, refs, bs, bd, tap4, planar) in enumerate(C.MCTF_APPLY_CASES):
        case = C.mctf_apply_case(seed, W, H, 24, refs, bs, bd)
        assert np.array_equal(impls.mctf_apply_expected(O, 'orc', case, tap4, planar), golden_mctf_apply['apply_%d' % k]), seed
    plane = golden_mctf_apply['var_plane']
    for (x, y, w, h), e in zip(golden_mctf_apply['var_blocks'], golden_mctf_apply['var_expect']):
        blk = np.ascontiguousarray(plane[y:y + h, x:x + w])
        assert O.orc_mctf_calc_var(P(blk), int(w), int(w), int(h)) == e


def test_oracle_frac_grid_golden(golden_frac):
    """two-pass 8-tap luma interpolation at every quarter-pel offset + SAD / SATD (xPatternRefinement's filtered blocks)"""
    from _libs import oracle, P, PO
    O = oracle()
    for ci, (seed, bd) in enumerate(C.FRAC_CASES):
        case = C.frac_case(seed, bit_depth=bd)
        S = case['stride']; base = case['margin'] * S + case['margin']
        for li, (fam, w, h, b) in enumerate(case['lists']):
            t = np.zeros((len(b), 7, 7), dtype=np.uint32)
            rt, alt = C.frac_filter_of(li)
            O.orc_frac_cost_grid(PO(case['org'], base), S, PO(case['ref'], base), S, P(np.ascontiguousarray(b)), len(b), fam, bd, rt, alt, P(t))
            assert np.array_equal(t, golden_frac['c%d_l%d' % (ci, li)]), (seed, fam, w, h)


def test_oracle_dep_quant_golden(golden_depquant):
    """DepQuant::xQuantDQ: the trellis restatement against levels / absSum / lastPos the reference produced (scalar and x86 members), rate tables from the
    reference's CABAC contexts; the Quantizer constants of initQuantBlock (double arithmetic) against the reference's"""
    import ctypes
    from _libs import dq_oracle, P
    O = dq_oracle()
    g = golden_depquant
    rows = C.dq_cases()
    assert np.array_equal(rows, g['cases'])
    nonzero = 0
    for i, row in enumerate(rows):
        w, h, bd, qp, lam1000, scale, decay10, mts, lf, sbt, intra, init_id, seed = [int(v) for v in row]
        coef = C.dq_inputs(row)
        k = np.zeros(9, dtype=np.int64)
        assert O.orc_dep_quant_constants(w, h, bd, qp, lam1000 / 1000.0, 8, P(k)) == 0
        assert np.array_equal(k, g['consts'][i]), (i, k, g['consts'][i])
        rates = np.ascontiguousarray(g['rates'][i])
        for scalar in (1, 0):
            q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
            assert O.orc_dep_quant(w, h, bd, qp, lam1000 / 1000.0, 8, C.dq_zero_out(row), lf, scalar, P(rates), P(coef), 1, P(q), ctypes.byref(s), ctypes.byref(l)) == 0
            name = 'q_x86_%d' % i
            want = g['q_scalar_%d' % i] if (scalar or name not in g) else g[name]
            assert np.array_equal(q, want), (i, scalar)
            assert (s.value, l.value) == tuple(int(v) for v in g['meta'][i, 0 if scalar else 1]), (i, scalar)
        nonzero += int(l.value >= 0)
    assert nonzero > 100


def test_oracle_dep_quant_chroma_golden(golden_depquant):
    """chroma components: context offsets of the chroma branch of xSetScanInfo, rate tables of the chroma context sets"""
    import ctypes
    from _libs import dq_oracle, P
    O = dq_oracle()
    g = golden_depquant
    rows = C.dq_chroma_cases()
    assert np.array_equal(rows, g['chroma_cases'])
    nz = 0
    for i, row in enumerate(rows):
        w, h, bd, qp, lam1000, scale, decay10, lf, intra, init_id, seed = [int(v) for v in row]
        coef = C.dq_chroma_inputs(row)
        rates = np.ascontiguousarray(g['chroma_rates'][i])
        for scalar in (0, 1):
            q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
            assert O.orc_dep_quant_chroma(w, h, bd, qp, lam1000 / 1000.0, 8, lf, scalar, P(rates), P(coef), 1, P(q), ctypes.byref(s), ctypes.byref(l)) == 0
            name = 'cq_scalar_%d' % i
            want = g[name] if (scalar and name in g) else g['cq_%d' % i]
            assert np.array_equal(q, want), (i, scalar)
        nz += int(l.value >= 0)
    assert nz > 30


def test_oracle_rdoq_golden(golden_rdoq):
    """QuantRDOQ2::xRateDistOptQuant (m_RDOQ == 2): the restatement (vvenc_b200/csrc/rdoq_core.h compiled for the CPU) against levels / absSum / lastPos the reference
    produced, fractional bits from the reference's CABAC contexts; luma / Cb / Cr, sign-bit hiding on and off, LFNST scan limit, SBT bin budget; the per-call
    constants (error scale in double arithmetic) against the reference's"""
    import ctypes
    from _libs import dq_oracle, P
    O = dq_oracle()
    g = golden_rdoq
    rows = C.rdoq_cases()
    assert np.array_equal(rows, g['cases'])
    nonzero = 0; hidden = 0
    for i, row in enumerate(rows):
        w, h, bd, qp, lam1000, scale, decay10, comp, lf, sbt, intra, sh, cb, thr, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_inputs(row)
        k = np.zeros(7, dtype=np.int32)
        assert O.orc_rdoq_constants(w, h, bd, qp, int(comp > 0), lf, sbt, thr, P(k)) == 0
        assert np.array_equal(k, g['consts'][i]), (i, k, g['consts'][i])
        rates = np.ascontiguousarray(g['rates'][i])
        q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
        assert O.orc_rdoq(w, h, bd, qp, int(comp > 0), lf, sbt, sh, lam1000 / 1000.0, thr, P(rates), P(coef), 1, P(q), ctypes.byref(s), ctypes.byref(l)) == 0
        assert np.array_equal(q, g['q_%d' % i]), (i, [int(v) for v in row])
        assert (s.value, l.value) == tuple(int(v) for v in g['meta'][i]), (i, [int(v) for v in row])
        q2 = np.zeros((h, w), dtype=np.int16); s2 = ctypes.c_int32(); l2 = ctypes.c_int32()                  # second engine
        assert O.orc_rdoq_v2(w, h, bd, qp, int(comp > 0), lf, sbt, sh, lam1000 / 1000.0, thr, P(rates), P(coef), 1, P(q2), ctypes.byref(s2), ctypes.byref(l2)) == 0
        assert np.array_equal(q2, g['q_%d' % i]) and (s2.value, l2.value) == tuple(int(v) for v in g['meta'][i]), ('engine 2', i, [int(v) for v in row])
        nonzero += int(l.value >= 0); hidden += int(sh and l.value >= 0)
    assert nonzero > 100 and hidden > 40


def test_oracle_rdoq_ts_golden(golden_rdoq):
    """QuantRDOQ::rateDistOptQuantTS (transform-skipped TUs): the restatement against levels / absSum the reference produced, fractional bits of the transform-skip context
    sets from the reference; the constants incl. the double-precision error scale against the reference's (equal to the last bit)"""
    import ctypes
    from _libs import dq_oracle, P
    O = dq_oracle()
    g = golden_rdoq
    rows = C.rdoq_ts_cases()
    assert np.array_equal(rows, g['ts_cases'])
    nonzero = 0
    for i, row in enumerate(rows):
        w, h, bd, qp, lam1000, amp, kind, comp, intra, delta, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_ts_inputs(row)
        k = np.zeros(3, dtype=np.int32); e = ctypes.c_double()
        assert O.orc_rdoq_ts_constants(w, h, bd, qp, delta, P(k), ctypes.byref(e)) == 0
        assert np.array_equal(k, g['ts_consts'][i]) and e.value == float(g['ts_err_scale'][i]), (i, k, g['ts_consts'][i])
        rates = np.ascontiguousarray(g['ts_rates'][i])
        q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32()
        assert O.orc_rdoq_ts(w, h, bd, qp, delta, lam1000 / 1000.0, P(rates), P(coef), 1, P(q), ctypes.byref(s)) == 0
        assert np.array_equal(q, g['tsq_%d' % i]) and s.value == int(g['ts_abs_sum'][i]), (i, [int(v) for v in row])
        nonzero += int(s.value > 0)
    assert nonzero > 80


def test_oracle_rdoq_bdpcm_golden(golden_rdoq):
    """QuantRDOQ::forwardRDPCM (BDPCM TUs): the restatement against the levels / absSum the reference produced on the transform-skip rows, direction 1 + (seed & 1)"""
    import ctypes
    from _libs import dq_oracle, P
    O = dq_oracle()
    g = golden_rdoq
    rows = C.rdoq_ts_cases()
    nonzero = 0
    for i, row in enumerate(rows):
        w, h, bd, qp, lam1000, amp, kind, comp, intra, delta, init_id, seed = [int(v) for v in row]
        coef = C.rdoq_ts_inputs(row)
        rates = np.ascontiguousarray(g['ts_rates'][i])
        q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32()
        assert O.orc_rdoq_bdpcm(w, h, bd, qp, delta, 1 + (seed & 1), lam1000 / 1000.0, P(rates), P(coef), 1, P(q), ctypes.byref(s)) == 0
        assert np.array_equal(q, g['bdq_%d' % i]) and s.value == int(g['bd_abs_sum'][i]), (i, [int(v) for v in row])
        nonzero += int(s.value > 0)
    assert nonzero > 70
