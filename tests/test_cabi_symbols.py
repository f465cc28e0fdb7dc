"""CPU-only: the C-ABI library is built, loads without a GPU, exports every symbol include/vvenc_b200.h declares,
and refuses to create a context when no CUDA device exists (no CPU fallback)."""
import os, re, subprocess, ctypes
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'vvenc_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(vvb_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    import vvenc_b200._lib as L
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = _declared()
    assert len(names) >= 25
    out = subprocess.check_output(['nm', '-D', '--defined-only', L.LIB_PATH], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if l.strip())
    missing = [n for n in names if n not in exported]
    assert missing == []
    # and the ctypes table binds the same set
    assert sorted(L.SYMBOLS) == names
    L.load()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import vvenc_b200 as V
    with pytest.raises(V.VvbError):
        V.CostEngine(0)


def test_product_never_imports_oracle():
    # the product package must not reference oracle/ (tests, smoke() and bench.py's cpu_baseline leg are the only users)
    pkg = os.path.join(ROOT, 'vvenc_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f), errors='ignore').read()
                assert 'liboracle' not in txt and 'refshim' not in txt and 'oracle/' not in txt.replace('oracle/vvc_tables.h', ''), f
                assert 'vvb_mock' not in txt, f
    # the oracle-backed mock of the C ABI (tests/mock) is for the host-logic tests only: neither the bench nor smoke() may touch it
    for f in ('bench.py', '__graft_entry__.py'):
        assert 'vvb_mock' not in open(os.path.join(ROOT, f)).read(), f


def test_reference_side_binding_resolves_only_declared_symbols():
    """integration/RdCostB200.h binds the library with dlsym: every name it asks for must be declared in the C ABI header (and hence exported)"""
    txt = open(os.path.join(ROOT, 'integration', 'RdCostB200.h')).read()
    txt_search = open(os.path.join(ROOT, 'integration', 'InterSearchB200.h')).read() + open(os.path.join(ROOT, 'integration', 'MCTFB200.h')).read() + \
        open(os.path.join(ROOT, 'integration', 'TrQuantB200.h')).read() + open(os.path.join(ROOT, 'integration', 'AffineGradientB200.h')).read()
    asked = sorted(set(re.findall(r'VVB_RESOLVE\(\s*\w+\s*,\s*(vvb_[a-z0-9_]+)\s*\)', txt + txt_search)))
    assert len(asked) >= 16 and 'vvb_affine_sobel' in asked and 'vvb_fwd_trquant' in asked and 'vvb_sad_search' in asked and 'vvb_frac_cost_grid' in asked and 'vvb_mctf_search_grid' in asked
    declared = set(_declared())
    assert [n for n in asked if n not in declared] == []
    # the trampolines cover every slot family the x86 back end overwrites (RdCostX86.h:3376-3425)
    for slot in ('DF_SSE', 'DF_SAD', 'DF_HAD', 'DF_HAD_fast', 'DF_HAD_2SAD', 'DF_SAD_WITH_MASK', 'm_afpDistortFuncX5', 'm_fxdWtdPredPtr'):
        assert slot in txt, slot
