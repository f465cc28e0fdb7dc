#!/usr/bin/env python3
"""Times vvb_rdoq_dev (QuantRDOQ2::xRateDistOptQuantFast on the device, one TU per thread) for a 2160p picture's worth of TUs per shape, CUDA events on the context
stream, and the CPU side on one thread: the reference's own member (oracle/_ref, incl. the probe's per-TU rig set-up) on a bounded sample, and the port (the same
text compiled by g++) on the whole list.  With a second argument (e.g. 1,4,16) the device is timed on that many pictures' worth of TUs per launch as well: the
kernel is bound by the serial chain of one TU, so the time per TU falls until the SMs are full.  usage: python tools/rq_bench.py [reps] [pictures,...] [only WxH]"""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import vvenc_b200 as V
    import vvenc_b200._lib as L
    from _libs import have_ref, refshim, dq_oracle, P
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    mult = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else []
    only = sys.argv[3] if len(sys.argv) > 3 else None
    eng = V.CostEngine(0)
    eng.set_rdoq_engine(int(os.environ.get('VVB_RDOQ_ENGINE', '1')))          # 2: accumulated templates + cost tables (not run on hardware yet)
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device('cuda', 0))
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_v6_rdoq.npz'))
    row0 = [i for i, r in enumerate(g['cases']) if int(r[7]) == 0][3]
    rates_flat = np.ascontiguousarray(g['rates'][row0])
    rates = eng.rdoq_rates(rates_flat)
    rs = np.random.RandomState(1)
    out = {}
    for (w, h) in ((8, 8), (16, 16), (32, 32), (64, 64)):
        if only and only != '%dx%d' % (w, h):
            continue
        n = (3840 // w) * (2160 // h)
        scale = rs.choice([3, 10, 40, 150, 600], size=(n, 1, 1))
        coef = rs.laplace(0, 1.0, size=(n, h, w)) * scale * (1.0 / (1 + np.add.outer(np.arange(h), np.arange(w))) ** 0.7)
        coef = np.clip(coef, -32768, 32767).astype(np.int32); coef[:, :, 32:] = 0; coef[:, 32:, :] = 0
        dcoef = torch.from_numpy(coef).cuda(); dq_ = torch.zeros((n, h, w), dtype=torch.int16, device='cuda')
        dsum = torch.zeros(n, dtype=torch.int32, device='cuda'); dlast = torch.zeros(n, dtype=torch.int32, device='cuda')
        par = eng.tu_par(w, h, 0, 0, 10, 32, sign_hiding=True); rqp = L.vvb_rdoq_par(57.3, 8, 0)
        cp = lambda t: ctypes.c_void_p(t.data_ptr())
        run = lambda: eng._chk(eng.lib.vvb_rdoq_dev(eng.h, ctypes.byref(par), ctypes.byref(rqp), ctypes.byref(rates), cp(dcoef), None, n, cp(dq_), cp(dsum), cp(dlast)))
        for _ in range(2):
            run()
        eng.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ext):
            e0.record(ext)
            for _ in range(reps):
                run()
            e1.record(ext)
        eng.synchronize(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        row = {'tus': n, 'ms_per_picture': ms, 'Mcoef_s': n * min(w, 32) * min(h, 32) / ms / 1e3, 'non_empty': int((dlast >= 0).sum().item())}
        O = dq_oracle()
        qq = np.zeros((n, h, w), dtype=np.int16); ss = np.zeros(n, dtype=np.int32); ll = np.zeros(n, dtype=np.int32)
        t0 = time.perf_counter()
        O.orc_rdoq(w, h, 10, 32, 0, 0, 0, 1, 57.3, 8, P(rates_flat), P(coef), n, P(qq), P(ss), P(ll))
        row['port_ms_per_picture_1thread'] = (time.perf_counter() - t0) * 1e3
        row['device_equals_port'] = bool(np.array_equal(dq_.cpu().numpy(), qq) and np.array_equal(dlast.cpu().numpy(), ll))
        if have_ref():
            R = refshim()
            m = min(n, max(20, 200000 // (w * h)))
            q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
            t0 = time.perf_counter()
            for i in range(m):
                R.refshim_rdoq(0, P(coef[i]), w, h, 10, 32, 0, 0, 0, 1, 0, 57.3, 8, 32, 0, P(q), ctypes.byref(s), ctypes.byref(l), None, None)
            row['reference_ms_per_picture_1thread_incl_rig_setup'] = (time.perf_counter() - t0) / m * n * 1e3
        for m in mult:                                   # m pictures' worth of TUs in one launch (the list repeated)
            if m <= 1 or n * m * w * h * 6 > 6e9:
                continue
            dc = dcoef.repeat(m, 1, 1); dqm = torch.zeros((n * m, h, w), dtype=torch.int16, device='cuda')
            ds = torch.zeros(n * m, dtype=torch.int32, device='cuda'); dl = torch.zeros(n * m, dtype=torch.int32, device='cuda')
            runm = lambda: eng._chk(eng.lib.vvb_rdoq_dev(eng.h, ctypes.byref(par), ctypes.byref(rqp), ctypes.byref(rates), cp(dc), None, n * m, cp(dqm), cp(ds), cp(dl)))
            runm(); eng.synchronize()
            with torch.cuda.stream(ext):
                e0.record(ext)
                for _ in range(max(1, reps // 2)):
                    runm()
                e1.record(ext)
            eng.synchronize(); torch.cuda.synchronize()
            row['ms_per_picture_at_%d_pictures' % m] = e0.elapsed_time(e1) / max(1, reps // 2) / m
            row['equal_at_%d_pictures' % m] = bool(torch.equal(dqm[(m - 1) * n:], dq_))
            del dc, dqm, ds, dl
        out['%dx%d' % (w, h)] = row
    print(json.dumps(out))


if __name__ == '__main__':
    main()
