"""Drop-in proof (-m gpu): the UNMODIFIED reference's RdCost function-pointer tables (CommonLib/RdCost.h:117-121) are patched
with trampolines into libvvenc_b200.so -- the `RdCost::_initRdCostB200()` of INTEGRATION.md section 2, compiled for real into
the reference probe (oracle/ref_shim.cpp: installB200) -- and the reference's own call sites are then driven through them:

  * DistParam + distFunc for every golden distortion case (SSE / SAD / HAD / HAD_fast / HAD_2SAD, strided, sub-sampled),
  * dmvrSadX5, the GEO mask SAD slot and m_fxdWtdPredPtr,
  * the xPatternSearch replay (InterSearch.cpp:2209-2251): reference loop + reference MV cost, SAD numbers from the GPU.

Results must equal the AVX2 table bit for bit.  Needs oracle/_ref (built where /root/reference exists; the .so travels)."""
import ctypes
import numpy as np
import pytest

import cases as C
import impls
from _libs import P, PO, have_ref, refshim

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason='oracle/_ref not built (reference sources absent at build time)')]


@pytest.fixture(scope='module')
def patched():
    import vvenc_b200._lib as VL
    L = refshim()
    L.refshim_b200_error.restype = ctypes.c_char_p
    L.refshim_b200_launches.restype = ctypes.c_uint64
    rc = L.refshim_install_b200(VL.LIB_PATH.encode())
    assert rc == 0, L.refshim_b200_error()
    return impls.RefImpl(opt=2)


def test_reference_rdcost_table_patched_with_b200(patched, golden):
    L = patched.L
    before = int(L.refshim_b200_launches())
    assert impls.run_dist(patched, golden['dist_rows'], golden['dist_expect']) == []
    launched = int(L.refshim_b200_launches()) - before
    # slots patched: base + log2(w) for w >= 2 (like initRdCostX86, the w == 1 slot keeps the scalar kernel); >10-bit uses table row [1] (RdCost.cpp:125-126)
    rows10 = sum(1 for r in golden['dist_rows'] if int(r[8]) <= 10 and int(r[1]) >= 2)
    assert launched >= rows10 > 0, (launched, rows10)


def test_reference_x5_mask_wsse_slots_patched(patched):
    L = patched.L
    rs = np.random.RandomState(606)
    for (w, h) in ((8, 8), (16, 16), (8, 16), (16, 8)):
        so = w + 16; sc = w + 24
        o = rs.randint(0, 1024, size=(h, so)).astype(np.int16); c = rs.randint(0, 1024, size=(h, sc)).astype(np.int16)
        for cc in (0, 1):
            a = np.zeros(5, dtype=np.uint64); b = np.zeros(5, dtype=np.uint64)
            L.refshim_sad_x5(1, PO(o, 0), so, PO(c, 8), sc, w, h, 10, 1, cc, P(a))
            L.refshim_sad_x5(2, PO(o, 0), so, PO(c, 8), sc, w, h, 10, 1, cc, P(b))
            assert np.array_equal(a, b), (w, h, cc)
    for (w, h) in ((4, 4), (8, 8), (64, 64), (128, 128)):
        o = rs.randint(0, 1024, size=(h, w + 8)).astype(np.int16); c = rs.randint(0, 1024, size=(h, w + 8)).astype(np.int16)
        wt = int(rs.randint(0, 1 << 17))
        assert L.refshim_fix_wsse(2, P(o), w + 8, P(c), w + 8, w, h, 10, wt) == L.refshim_fix_wsse(1, P(o), w + 8, P(c), w + 8, w, h, 10, wt)
    for (w, h) in ((8, 8), (16, 16), (32, 64), (64, 64)):
        for stepX in (1, -1):
            ms = w + 8
            o = rs.randint(0, 1024, size=(h, w + 8)).astype(np.int16); c = rs.randint(0, 1024, size=(h, w + 8)).astype(np.int16)
            msk = rs.randint(0, 2, size=(h + 2, ms)).astype(np.int16)
            start = 0 if stepX == 1 else w - 1
            a = L.refshim_sad_mask(1, P(o), w + 8, P(c), w + 8, w, h, PO(msk, start), ms, stepX, -w * stepX, 10, 0)
            b = L.refshim_sad_mask(2, P(o), w + 8, P(c), w + 8, w, h, PO(msk, start), ms, stepX, -w * stepX, 10, 0)
            assert a == b, (w, h, stepX)


def test_reference_pattern_search_loop_on_b200_sad(patched, golden):
    """the reference's full-search loop (raster order, strict '<', MV cost from RdCost) with the SAD pointer patched"""
    sc = C.search_case()
    keep = [i for i, b in enumerate(sc['blk']) if (b[5] - b[4] + 1) * (b[7] - b[6] + 1) <= 17 * 17]    # per-call launches: keep the replay short
    sub = dict(sc); sub['blk'] = np.ascontiguousarray(sc['blk'][keep])
    for ss in (0, 1):
        got = patched.full_search(sub, ss)
        assert np.array_equal(got, golden['search_best_ss%d' % ss][keep]), ss
