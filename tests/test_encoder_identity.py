"""Whole-encoder bitstream identity, the reference's own criterion for a back end (cmake/modules/vvencTests.cmake:52-53: --SIMD=SCALAR vs default must give the
same .vvc).  oracle/_ref/enc_identity drives the UNMODIFIED reference encoder through its public API; with a library path it installs integration/RdCostB200.h's
and AffineGradientB200.h's tables in every RdCost / AffineGradientSearch the encoder creates (linker --wrap of the two x86 init calls, no source change).

  CPU (here)  : the tables call the oracle-backed mock of the C ABI -> pins the binding + the oracle's arithmetic against the AVX2 encoder, bitstream for bitstream
  GPU (-m gpu): the tables call libvvenc_b200.so, one launch per xGetSAD / xGetSSE / xGetHADs call -> the CUDA kernels under the real encoder control flow"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, 'oracle', '_ref', 'enc_identity')
MOCK = os.path.join(ROOT, 'tests', 'mock', '_build', 'libvvb_mock.so')
pytestmark = pytest.mark.skipif(not os.path.exists(BIN), reason='oracle/_ref/enc_identity not built (needs /root/reference at build time)')


def _encode(tmp_path, clip, W, H, F, preset, qp, lib=None, timeout=600, tu=False, rdoq=False, mctf=False, ts=False, bdpcm=False):
    out = str(tmp_path / ('b200.vvc' if lib else 'avx2.vvc'))
    cmd = [BIN, clip, str(W), str(H), str(F), str(preset), str(qp), out] + ([lib] if lib else []) + ((['all'] if mctf else ['turdoq'] if rdoq else ['tu']) if tu else [])
    env = dict(os.environ)
    if ts:
        env['VVB_ENC_TS'] = '1'              # transform skip tried on every eligible TU (the presets leave it to the screen-content detector); both arms set it
    if bdpcm:
        env['VVB_ENC_BDPCM'] = '1'           # ... and block DPCM next to it
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-1500:])
    line = [l for l in r.stdout.splitlines() if l.startswith('ENC ')][-1]
    kv = dict(f.split('=') for f in line.split()[1:])
    return open(out, 'rb').read(), kv


def _identity_tu(tmp_path, W, H, F, preset, qp, lib, want_dq, timeout=900):
    """the same with TrQuant::transformNxN / invTransformNxN routed through the library (forward transforms, transform skip, LFNST, the DepQuant trellis with rate
    tables from the live CABAC contexts, the inverse path with either dequantiser and the inverse LFNST)"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from _clips import write_clip
    clip = str(tmp_path / 'clip.yuv')
    write_clip(clip, W, H, F, seed=W + F)
    a, ka = _encode(tmp_path, clip, W, H, F, preset, qp)
    b, kb = _encode(tmp_path, clip, W, H, F, preset, qp, lib, timeout, tu=True)
    assert int(kb['tu_fwd']) > 1000 and int(ka['tu_fwd']) == 0
    if want_dq:
        assert int(kb['tu_dq']) > 1000, kb                 # the preset enables dependent quantisation: the trellis ran in the library
    assert int(kb['tu_inv']) > 500, kb                     # the inverse path (plain or DepQuant dequantiser) ran in the library
    if want_dq:
        assert int(kb['tu_inv_lfnst']) > 500, kb           # these presets enable LFNST: the inverse LFNST ran in the library too
    assert len(a) > 200 and a == b, (len(a), len(b), ka, kb)
    return kb


def _identity_rdoq(tmp_path, W, H, F, preset, qp, lib, timeout=900):
    """the TU seam as above plus QuantRDOQ2::xRateDistOptQuant (the fast RDOQ of the presets faster / fast) through vvb_rdoq: fractional bits of the live CABAC
    contexts, the member's last-position table, the coded-block-flag context of the live CU -- level decisions in the library"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from _clips import write_clip
    clip = str(tmp_path / 'clip.yuv')
    write_clip(clip, W, H, F, seed=W + F)
    a, ka = _encode(tmp_path, clip, W, H, F, preset, qp)
    b, kb = _encode(tmp_path, clip, W, H, F, preset, qp, lib, timeout, tu=True, rdoq=True)
    assert int(kb['tu_fwd']) > 1000 and int(kb['tu_rdoq']) > 900 and int(ka['tu_rdoq']) == 0, kb
    assert len(a) > 200 and a == b, (len(a), len(b), ka, kb)
    return kb


def _identity_widest(tmp_path, W, H, F, preset, qp, lib, timeout=900):
    """`turdoq` on presets with dependent quantisation and LFNST: on top of the TU seam above, LFNST on the chroma TUs of the separate tree of I-slices (kernel set from
    the chroma mode or the co-located luma mode) and on ISP luma TUs, the chroma TUs of single-tree LFNST CUs (zero-out without the kernel), joint Cb-Cr TUs -- what
    stays with the members is TUs with a side below 4 and little else"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from _clips import write_clip
    clip = str(tmp_path / 'clip.yuv')
    write_clip(clip, W, H, F, seed=W + F)
    a, ka = _encode(tmp_path, clip, W, H, F, preset, qp)
    b, kb = _encode(tmp_path, clip, W, H, F, preset, qp, lib, timeout, tu=True, rdoq=True)
    assert int(kb['tu_fwd']) > 15000 and int(kb['tu_dq']) > 15000 and int(kb['tu_inv_lfnst']) > 5000, kb
    assert int(kb['tu_ref']) * 4 < int(kb['tu_fwd']), kb               # the narrow routing leaves more TUs to the members than it takes (tu_ref > 2 x tu_fwd)
    assert len(a) > 200 and a == b, (len(a), len(b), ka, kb)
    return kb


def _identity_mctf(tmp_path, W, H, F, preset, qp, lib, timeout=900):
    """`all`: everything above plus the block-matching errors of the MCTF pre-analysis -- MCTF::initMCTF_X86 wrapped, the error pointers and m_calcVar answered per call by
    the library (integration/MCTFB200.h: installB200( MCTF& )) under the unmodified MCTF::motionEstimationLuma control.  Nine frames, so that the temporal filter has
    its neighbours; a perturbed error changes the bitstream (checked when the test was written), so identity pins the errors"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from _clips import write_clip
    clip = str(tmp_path / 'clip.yuv')
    write_clip(clip, W, H, F, seed=W + F)
    a, ka = _encode(tmp_path, clip, W, H, F, preset, qp)
    b, kb = _encode(tmp_path, clip, W, H, F, preset, qp, lib, timeout, tu=True, rdoq=True, mctf=True)
    assert int(kb['mctf_installs']) >= 1 and int(kb['mctf_calls']) > 10000 and int(ka['mctf_calls']) == 0, kb
    assert int(kb['tu_fwd']) > 5000 and int(kb['dist_calls']) > 10000, kb
    assert len(a) > 200 and a == b, (len(a), len(b), ka, kb)
    return kb


def _identity_ts(tmp_path, W, H, F, preset, qp, lib, min_ts, timeout=900):
    """screen-content style clip with transform skip enabled in both arms: the transform-skipped TUs go through xTransformSkip + rateDistOptQuantTSB200 (-> vvb_rdoq_ts)
    and the inverse path of skipped transforms, next to everything `turdoq` routes"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from _clips import write_clip_scc
    clip = str(tmp_path / 'clip.yuv')
    write_clip_scc(clip, W, H, F, seed=W + F)
    a, ka = _encode(tmp_path, clip, W, H, F, preset, qp, ts=True)
    b, kb = _encode(tmp_path, clip, W, H, F, preset, qp, lib, timeout, tu=True, rdoq=True, ts=True)
    assert int(kb['tu_rdoq_ts']) >= min_ts and int(ka['tu_rdoq_ts']) == 0 and int(kb['tu_fwd']) > 1000, kb
    assert len(a) > 200 and a == b, (len(a), len(b), ka, kb)
    return kb


def _identity_bdpcm(tmp_path, W, H, F, preset, qp, lib, min_bdpcm, timeout=900):
    """as _identity_ts with block DPCM enabled as well: BDPCM TUs go through xTransformSkip + forwardRDPCMB200 (-> vvb_rdoq_bdpcm) and, on the inverse side, through
    invTransformNxNB200 (the running sums of Quant::dequant on the host, then the library's inverse of skipped transforms)"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from _clips import write_clip_scc
    clip = str(tmp_path / 'clip.yuv')
    write_clip_scc(clip, W, H, F, seed=W + F)
    a, ka = _encode(tmp_path, clip, W, H, F, preset, qp, ts=True, bdpcm=True)
    b, kb = _encode(tmp_path, clip, W, H, F, preset, qp, lib, timeout, tu=True, rdoq=True, ts=True, bdpcm=True)
    assert int(kb['tu_bdpcm']) >= min_bdpcm and int(ka['tu_bdpcm']) == 0 and int(kb['tu_rdoq_ts']) > 40, kb
    assert len(a) > 200 and a == b, (len(a), len(b), ka, kb)
    return kb


def _identity(tmp_path, W, H, F, preset, qp, lib, timeout=600):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from _clips import write_clip
    clip = str(tmp_path / 'clip.yuv')
    write_clip(clip, W, H, F, seed=W + F)
    a, ka = _encode(tmp_path, clip, W, H, F, preset, qp)
    b, kb = _encode(tmp_path, clip, W, H, F, preset, qp, lib, timeout)
    assert int(kb['rdcost_installs']) >= 1 and int(kb['dist_calls']) > 1000 and int(ka['dist_calls']) == 0
    assert len(a) > 200 and a == b, (len(a), len(b), ka, kb)
    return kb


@pytest.mark.parametrize("W,H,F,preset,qp", [(80, 44, 4, 0, 37), (80, 44, 4, 2, 37), (176, 144, 3, 1, 32)])
def test_bitstream_identity_with_b200_tables_on_the_oracle(tmp_path, W, H, F, preset, qp):
    if not os.path.exists(MOCK):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    _identity(tmp_path, W, H, F, preset, qp, MOCK)


@pytest.mark.parametrize("W,H,F,preset,qp,dq", [(80, 44, 4, 0, 37, False), (80, 44, 3, 2, 37, True), (176, 144, 2, 1, 32, True)])
def test_bitstream_identity_with_the_tu_seam_on_the_oracle(tmp_path, W, H, F, preset, qp, dq):
    if not os.path.exists(MOCK):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    _identity_tu(tmp_path, W, H, F, preset, qp, MOCK, dq)


@pytest.mark.parametrize("W,H,F,preset,qp", [(80, 44, 4, 0, 37), (176, 144, 3, 0, 27), (416, 240, 8, 0, 37)])
def test_bitstream_identity_with_the_rdoq_seam_on_the_oracle(tmp_path, W, H, F, preset, qp):
    """preset faster runs Quant::m_RDOQ == 2 with sign-bit hiding and selective RDOQ (vvencCfg.cpp:2675-2677): 1 000 / 7 000 / 16 500 TUs through xRateDistOptQuantB200;
    (416, 240, 8 frames, faster, QP 37) is BASELINE configs[0]"""
    if not os.path.exists(MOCK):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    _identity_rdoq(tmp_path, W, H, F, preset, qp, MOCK)


@pytest.mark.parametrize("W,H,F,preset,qp", [(80, 44, 3, 2, 37), (176, 144, 2, 1, 32), (176, 144, 2, 3, 27)])
def test_bitstream_identity_with_the_widest_tu_seam_on_the_oracle(tmp_path, W, H, F, preset, qp):
    """presets medium / fast / slow: 19 000 / 57 000 / 344 000 forward TUs through the library, 3 600 / 1 600 / 68 000 left to the members"""
    if not os.path.exists(MOCK):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    _identity_widest(tmp_path, W, H, F, preset, qp, MOCK)


@pytest.mark.parametrize("W,H,F,preset,qp", [(80, 44, 9, 2, 37), (176, 144, 9, 0, 32), (176, 144, 9, 1, 32)])
def test_bitstream_identity_with_the_mctf_errors_on_the_oracle(tmp_path, W, H, F, preset, qp):
    """16 700 / 44 900 / 67 300 MCTF error calls through the library next to the distortion tables and the widest TU seam"""
    if not os.path.exists(MOCK):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    _identity_mctf(tmp_path, W, H, F, preset, qp, MOCK)


@pytest.mark.parametrize("W,H,F,preset,qp,min_ts", [(80, 44, 4, 0, 32, 40), (176, 144, 3, 0, 27, 1000), (176, 144, 2, 1, 27, 3000)])
def test_bitstream_identity_with_transform_skip_rdoq_on_the_oracle(tmp_path, W, H, F, preset, qp, min_ts):
    """50 / 1 300 / 3 400 transform-skipped TUs through QuantRDOQ::rateDistOptQuantTS's replacement (presets faster and fast: with RDOQ 2 and with dependent quantisation)"""
    if not os.path.exists(MOCK):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    _identity_ts(tmp_path, W, H, F, preset, qp, MOCK, min_ts)


@pytest.mark.parametrize("W,H,F,preset,qp,min_bdpcm", [(80, 44, 4, 0, 32, 100), (176, 144, 3, 0, 27, 2000), (176, 144, 2, 1, 27, 6000), (176, 144, 2, 2, 32, 7000)])
def test_bitstream_identity_with_bdpcm_on_the_oracle(tmp_path, W, H, F, preset, qp, min_bdpcm):
    """124 / 2 700 / 8 200 / 9 400 BDPCM TUs (forward and inverse) next to 50 - 14 000 plain transform-skipped ones: presets faster, fast, medium"""
    if not os.path.exists(MOCK):
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'tests', 'mock')])
    _identity_bdpcm(tmp_path, W, H, F, preset, qp, MOCK, min_bdpcm)


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,F,preset,qp,dq", [(80, 44, 4, 0, 37, False), (80, 44, 3, 2, 37, True), (176, 144, 2, 1, 32, True)])
def test_bitstream_identity_with_the_tu_seam_on_the_gpu(tmp_path, W, H, F, preset, qp, dq):
    import vvenc_b200._lib as VL
    kb = _identity_tu(tmp_path, W, H, F, preset, qp, VL.LIB_PATH, dq, timeout=1500)
    print('encoder identity with the TU seam on the GPU:', W, H, F, preset, kb)


@pytest.mark.gpu
@pytest.mark.parametrize("W,H,F,preset,qp", [(80, 44, 4, 0, 37), (80, 44, 3, 2, 37), (416, 240, 8, 0, 37)])
def test_bitstream_identity_with_b200_tables_on_the_gpu(tmp_path, W, H, F, preset, qp):
    """(416, 240, 8 frames, faster, QP 37) is BASELINE configs[0]"""
    import vvenc_b200._lib as VL
    kb = _identity(tmp_path, W, H, F, preset, qp, VL.LIB_PATH, timeout=1500)
    print('encoder identity on the GPU:', W, H, F, preset, kb)
