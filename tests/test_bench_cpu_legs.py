"""bench.py's CPU legs (the reference arm of the step and the per-row baselines) run on the host only: keep them runnable here, with a tiny budget."""
import os
import sys

import pytest

from _libs import have_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.skipif(not have_ref(), reason='oracle/_ref not built')


def test_cpu_rows_report_every_widened_row():
    import bench
    r = bench.cpu_rows(threads=2, budget_s=0.02)
    assert r['kind'] == 'reference' and r['cores'] == 2
    for n in bench.SIZES:
        assert r['tu_roundtrip'][str(n)]['tu_per_s'] > 0
    assert r['mctf_match_16x16']['cand_per_s'] > 0
    for n in (8, 16, 32):
        assert r['frac_satd_grid'][str(n)]['cand_per_s'] > 0
    assert r['mctf_apply']['pels_per_s'] > 0


def test_cpu_arm_reports_the_step_metric():
    import bench
    r = bench.cpu_arm(sample_budget_s=0.3, threads=2, quiet=True)
    assert r['unit'] == 'candidate-blocks/s' and r['value'] > 0 and r['kind'] == 'reference'
    assert set(r['legs']) == set(bench.SIZES)
