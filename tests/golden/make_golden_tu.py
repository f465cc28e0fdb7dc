#!/usr/bin/env python3
"""Regenerates tests/golden/golden_v2_tu.npz (inverse path + TU round trip) from the UNMODIFIED reference (oracle/_ref, built from
/root/reference by oracle/Makefile.ref).  Run in the build container only:  python tests/golden/make_golden_tu.py
Every expected value comes from the reference's AVX2 path and is cross-checked against its scalar path."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases as C
import impls


def main():
    ref = [impls.RefImpl(0), impls.RefImpl(1)]
    out = {}
    rows = C.itq_cases(); coefs = []; resis = []
    for row in rows:
        th, tv, w, h, st, kind, qp, bd, seed = [int(v) for v in row]
        q = C.itq_inputs(row)
        r = [x.inv_transform_quant(th, tv, q, w, h, bd, qp, st) for x in ref]
        assert np.array_equal(r[0][0], r[1][0]) and np.array_equal(r[0][1], r[1][1]), ('scalar != AVX2', row)
        coefs.append(r[1][0].reshape(-1)); resis.append(r[1][1].reshape(-1))
    out['itq_rows'] = rows; out['itq_coef'] = np.concatenate(coefs); out['itq_resi'] = np.concatenate(resis)
    rows = C.rt_cases(); qs = []; recos = []; meta = np.zeros((len(rows), 5), dtype=np.int64)
    for i, row in enumerate(rows):
        th, tv, w, h, so, ps, amp, qp, irap, bd, seed = [int(v) for v in row]
        org, pred = C.rt_inputs(row)
        r = [x.tu_roundtrip(th, tv, org, so, pred, ps, w, h, bd, qp, irap) for x in ref]
        assert np.array_equal(r[0][0], r[1][0]) and np.array_equal(r[0][1], r[1][1]) and r[0][2] == r[1][2], ('scalar != AVX2', row)
        qs.append(r[1][0].reshape(-1)); recos.append(r[1][1].reshape(-1)); meta[i] = r[1][2]
    out['rt_rows'] = rows; out['rt_q'] = np.concatenate(qs); out['rt_reco'] = np.concatenate(recos); out['rt_meta'] = meta
    path = os.path.join(HERE, 'golden_v2_tu.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, {k: v.shape for k, v in out.items()}, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
