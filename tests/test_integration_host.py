"""Host logic of the reference-side bindings in integration/ (RdCostB200.h, InterSearchB200.h), on the CPU: the bindings are compiled into the reference
probe (oracle/_ref) and bound to tests/mock -- the C ABI answered by the oracle -- so that what is under test is the glue itself: argument marshalling
out of DistParam / TZSearchStruct / RdCost, and the replay of the reference's selection rounds on the returned tables.  Checked against the reference's
own code in the same process:

  * RdCost tables patched by installB200() vs the AVX2 table on every golden distortion row,
  * AffineGradientSearch pointers patched by installB200() vs the AVX2 kernels on every affine case row,
  * xPatternSearchB200 and B200RowSearch vs InterSearch::xPatternSearch (member call), all subShift modes, two AMVR shifts,
  * xTZSearchB200 vs InterSearch::xTZSearch (member call): the unmodified member walks the dense SAD table of one vvb_sad_search launch -- diamond / enhanced /
    fast settings, integer early termination, first-search stop; with a reach too small for the walk the per-block path answers the rest; per PU and
    per row (B200RowSearch: one launch per block size fills all tables, then the walks), and from four worker threads at once (a context and a table each),
  * xPatternSearchFracDIFB200 vs InterSearch::xPatternSearchFracDIF (member call), 8/6/4-tap ME filters, SATD and SAD, alt half-pel, square and rectangular,
  * motionEstimationLumaB200 vs MCTF::motionEstimationLuma (member call): first level, chained level and the doubleRes final level, search patterns 0/1/2,
    6- and 4-tap search filters, pictures with partial border blocks,
  * bilateralFilterB200 vs MCTF::bilateralFilter (member call) on whole luma pictures: units 8/16/32, 2..8 neighbour pictures, both filter sets, QP on both
    sides of the planar-correction threshold; and on 4:2:0 pictures, all three components (chroma through the same entry point: half-size units, vectors pre-shifted),
  * xTQuantB200 / invTransformNxNB200 vs TrQuant::xT + Quant::quant (+ xNeedRDOQ) / Quant::dequant + xIT on a TransformUnit: every row of the parity tables.

The same bindings run against libvvenc_b200.so in tests/test_gpu_dropin.py (-m gpu)."""
import json
import os
import subprocess
import sys

import pytest

from _libs import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref not built')


@pytest.fixture(scope='module')
def result():
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    mock = os.path.join(ROOT, 'tests', 'mock', '_build', 'libvvb_mock.so')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_integration_host_run.py'), mock], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith('RESULT ')]
    assert line, out.stdout[-2000:]
    return json.loads(line[-1][len('RESULT '):])


def test_rdcost_tables_patched_through_the_binding(result):
    assert result['dist_rows'] > 600 and result['dist_mismatches'] == 0
    assert result['affine']['cases'] >= 12 and result['affine']['bad'] == 0


def test_pattern_search_binding_equals_the_member(result):
    assert len(result['search']) == 12
    for r in result['search']:
        assert r['rc'] == [0, 0], r
        assert r['member_eq_b200'] and r['member_eq_rows'], r


def test_fractional_search_binding_equals_the_member(result):
    assert len(result["frac"]) >= 200 and any(r["fsp"] == 1 and r["w"] != r["h"] for r in result["frac"]) and any(r["had"] == 2 for r in result["frac"])
    for r in result['frac']:
        assert r['rc'] == 0 and r['eq'], r


def test_mctf_search_binding_equals_the_member(result):
    assert len(result['mctf']) == 24
    for r in result['mctf']:
        assert r['rc'] == [0] * 6 and all(r['eq']), r
        assert r['moving'] > r['blocks'] // 2, r                 # the synthetic displacement is found: the fields are not trivially zero


def test_transform_quant_binding_equals_the_members(result):
    assert result['tu_fwd']['cases'] > 250 and result['tu_fwd']['bad'] == []
    assert result['tu_fwd_lfnst']['cases'] == 288 and result['tu_fwd_lfnst']['bad'] == []      # LFNST forward
    assert result['tu_fwd_sdh']['cases'] > 250 and result['tu_fwd_sdh']['bad'] == [] and result['tu_fwd_sdh']['levels_changed_by_hiding'] > 60      # sign-bit hiding on
    assert result['tu_inv']['cases'] > 250 and result['tu_inv']['bad'] == []
    assert result['tu_ts_chroma']['cases'] == 220 and result['tu_ts_chroma']['inverse_cases'] > 20 and result['tu_ts_chroma']['bad'] == []      # transform skip, chroma components
    assert result['tu_inv_dq']['cases'] == 168 and result['tu_inv_dq']['bad'] == []      # DepQuant dequantiser
    assert result['tu_inv_lfnst']['cases'] == 288 and result['tu_inv_lfnst']['bad'] == []   # inverse LFNST
    assert result['dep_quant_chroma']['cases'] == 72 and result['dep_quant_chroma']['bad'] == []
    assert result['dep_quant']['cases'] == 216 and result['dep_quant']['non_empty'] > 100 and result['dep_quant']['bad'] == []      # dependent quantisation (xQuantDQB200)


def test_mctf_apply_binding_equals_the_member(result):
    assert len(result['mctf_apply']) == 10
    for r in result['mctf_apply']:
        assert r['rc'] == 0 and r['eq'] and r['changed'], r
    assert len(result['mctf_apply420']) == 8
    for r in result['mctf_apply420']:
        assert r['rc'] == [0, 0] and r['eq_luma'] and r['eq_chroma'] and r['chroma_changed'], r


def test_tz_search_binding_equals_the_member(result):
    assert len(result['tz']) == 12
    for r in result['tz']:
        assert r['rc'] == [0, 0] and r['eq'], r
        assert r['moving'] > r['blocks'] // 2 and r['hits'] > 20 * r['blocks'] and r['row_hits'] == r['hits'] and r['row_misses'] == r['misses'], r
    assert all(r['misses'] == 0 for r in result['tz'] if r['cfg'][6] == 80)               # the window guess covers the whole walk
    assert all(r['misses'] > 0 for r in result['tz'] if r['cfg'][6] == 14)                # ... and the fallback is exercised when it cannot


def test_rdoq_binding_equals_the_member():
    """xRateDistOptQuantB200 vs QuantRDOQ2::xRateDistOptQuant (member call) on the TU rig, bound to the oracle-backed mock: every row of cases.rdoq_cases()"""
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    mock = os.path.join(ROOT, 'tests', 'mock', '_build', 'libvvb_mock.so')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', '_integration_host_run.py'), mock, 'rdoq'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith('RESULT ')][-1][len('RESULT '):])
    r = res['rdoq']
    assert r['cases'] == 224 and r['non_empty'] > 100 and r['non_empty_with_hiding'] > 40 and r['bad'] == [], r
    t = res['rdoq_ts']                          # rateDistOptQuantTSB200 vs QuantRDOQ::rateDistOptQuantTS
    assert t['cases'] == 140 and t['non_empty'] > 80 and t['bad'] == [], t
    b = res['rdoq_bdpcm']                       # forwardRDPCMB200 vs QuantRDOQ::forwardRDPCM, and the inverse path of the BDPCM levels (invTransformNxNB200 vs the member)
    assert b['cases'] == 140 and b['non_empty'] > 70 and b['bad'] == [], b
