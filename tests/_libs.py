"""ctypes loaders shared by the tests: the CPU oracle (oracle/_build/liboracle.so), the optional reference
probe (oracle/_ref/libvvenc_refshim.so, only where /root/reference was available to build it) and the product
C-ABI library (vvenc_b200/csrc/libvvenc_b200.so)."""
import ctypes, os, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
c_i16p = ctypes.c_void_p


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def PO(a, off):
    """pointer to element offset `off` (may be inside a margin) of a contiguous array"""
    return ctypes.c_void_p(a.ctypes.data + off * a.itemsize)


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        so = os.path.join(ROOT, 'oracle', '_build', 'liboracle.so')
        if not os.path.exists(so):
            subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
        L = ctypes.CDLL(so)
        for name in ('orc_sad', 'orc_sse', 'orc_had', 'orc_had2sad', 'orc_dist', 'orc_sad_mask', 'orc_fix_wsse', 'orc_mv_cost'):
            getattr(L, name).restype = ctypes.c_uint64
        L.orc_mv_bits.restype = ctypes.c_uint32
        L.orc_mv_cost.argtypes = [ctypes.c_double] + [ctypes.c_int] * 6
        L.orc_full_search.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _oracle = L
    return _oracle


_dqoracle = None


def dq_oracle():
    """oracle/_build/libdqoracle.so: the dependent-quantisation restatement (vvenc_b200/csrc/depquant_core.h) compiled for the CPU"""
    global _dqoracle
    if _dqoracle is None:
        so = os.path.join(ROOT, 'oracle', '_build', 'libdqoracle.so')
        if not os.path.exists(so):
            subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
        L = ctypes.CDLL(so)
        L.orc_dep_quant.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double] + [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_dep_quant_chroma.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double] + [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_dep_quant_constants.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_int, ctypes.c_void_p]
        # oracle/rdoq_oracle.cpp (vvenc_b200/csrc/rdoq_core.h compiled for the CPU) lives in the same library
        L.orc_rdoq.argtypes = [ctypes.c_int] * 8 + [ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_rdoq_constants.argtypes = [ctypes.c_int] * 8 + [ctypes.c_void_p]
        L.orc_rdoq_v2.argtypes = L.orc_rdoq.argtypes
        L.orc_rdoq_ts.argtypes = [ctypes.c_int] * 5 + [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.orc_rdoq_ts_constants.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p]
        L.orc_rdoq_bdpcm.argtypes = [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _dqoracle = L
    return _dqoracle


_ref = None


def refshim_path():
    return os.path.join(ROOT, 'oracle', '_ref', 'libvvenc_refshim.so')


def have_ref():
    return os.path.exists(refshim_path())


def refshim():
    global _ref
    if _ref is None:
        L = ctypes.CDLL(refshim_path())
        for name in ('refshim_dist', 'refshim_sad_mask', 'refshim_fix_wsse', 'refshim_mv_cost'):
            getattr(L, name).restype = ctypes.c_uint64
        L.refshim_mv_bits.restype = ctypes.c_uint32
        L.refshim_mv_cost.argtypes = [ctypes.c_double] + [ctypes.c_int] * 6
        L.refshim_set_simd.restype = ctypes.c_char_p
        L.refshim_full_search.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                          ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.refshim_pattern_search_member.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.refshim_dep_quant.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_double] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 5
        L.refshim_dep_quant_comp.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_double] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 5
        L.refshim_dep_quant_b200_comp.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_double] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
        L.refshim_dep_quant_b200.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 8 + [ctypes.c_double] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
        L.refshim_rdoq.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 9 + [ctypes.c_double] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5
        L.refshim_rdoq_b200.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 9 + [ctypes.c_double] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
        L.refshim_rdoq_ts.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 5
        L.refshim_rdoq_ts_b200.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 2
        L.refshim_rdoq_bdpcm.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 7 + [ctypes.c_double, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
        L.refshim_rdoq_bdpcm_b200.argtypes = [ctypes.c_int, ctypes.c_void_p] + [ctypes.c_int] * 6 + [ctypes.c_double, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 4
        L.refshim_set_simd(b'AVX2')
        _ref = L
    return _ref
