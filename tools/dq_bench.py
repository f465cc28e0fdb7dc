#!/usr/bin/env python3
"""Times vvb_dep_quant_dev (DepQuant::xQuantDQ trellis on the device) for a 2160p picture's worth of TUs per shape, CUDA events on the context stream, and the
reference's own DepQuant::xQuantDQ (oracle/_ref, one thread) on a bounded sample of the same TUs.  usage: python tools/dq_bench.py [reps]"""
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
    eng = V.CostEngine(0)
    eng.set_depquant_engine(int(os.environ.get('VVB_DQ_ENGINE', '1')))
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device('cuda', 0))
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_v5_depquant.npz'))
    rates_flat = np.ascontiguousarray(g['rates'][3])
    rates = eng.dq_rates(rates_flat)
    rs = np.random.RandomState(1)
    out = {}
    for (w, h) in ((8, 8), (16, 16), (32, 32), (64, 64)):
        n = (3840 // w) * (2160 // h)
        scale = rs.choice([3, 10, 40, 150, 600], size=(n, 1, 1))
        coef = rs.laplace(0, 1.0, size=(n, h, w)) * scale * (1.0 / (1 + np.add.outer(np.arange(h), np.arange(w))) ** 0.7)
        coef = np.clip(coef, -32768, 32767).astype(np.int32); coef[:, :, 32:] = 0; coef[:, 32:, :] = 0
        dcoef = torch.from_numpy(coef).cuda(); dq_ = torch.zeros((n, h, w), dtype=torch.int16, device='cuda')
        dsum = torch.zeros(n, dtype=torch.int32, device='cuda'); dlast = torch.zeros(n, dtype=torch.int32, device='cuda')
        par = eng.tu_par(w, h, 0, 0, 10, 32); dqp = L.vvb_dq_par(57.3, 8, 0, 0, 0)
        cp = lambda t: ctypes.c_void_p(t.data_ptr())
        run = lambda: eng._chk(eng.lib.vvb_dep_quant_dev(eng.h, ctypes.byref(par), ctypes.byref(dqp), ctypes.byref(rates), cp(dcoef), None, n, cp(dq_), cp(dsum), cp(dlast)))
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
        # CPU: the reference's member on a bounded sample (one thread), else the port
        m = min(n, max(20, 200000 // (w * h)))
        q = np.zeros((h, w), dtype=np.int16); s = ctypes.c_int32(); l = ctypes.c_int32()
        if have_ref():
            R = refshim()
            t0 = time.perf_counter()
            for i in range(m):
                R.refshim_dep_quant(P(coef[i]), w, h, 10, 32, 0, 0, 0, 0, 57.3, 8, 1, 32, 0, P(q), ctypes.byref(s), ctypes.byref(l), None, None)
            dt = time.perf_counter() - t0
            row['cpu_kind'] = 'reference (x86 members, 1 thread, incl. per-TU rig set-up)'
        else:
            O = dq_oracle()
            qq = np.zeros((m, h, w), dtype=np.int16); ss = np.zeros(m, dtype=np.int32); ll = np.zeros(m, dtype=np.int32)
            t0 = time.perf_counter()
            O.orc_dep_quant(w, h, 10, 32, 57.3, 8, 0, 0, 0, P(rates_flat), P(coef[:m]), m, P(qq), P(ss), P(ll))
            dt = time.perf_counter() - t0
            row['cpu_kind'] = 'port (1 thread)'
        row['cpu_ms_per_picture_1thread'] = dt / m * n * 1e3
        row['gpu_over_cpu_thread'] = row['cpu_ms_per_picture_1thread'] / ms
        out['%dx%d' % (w, h)] = row
    print(json.dumps(out))


if __name__ == '__main__':
    main()
