#!/usr/bin/env python3
"""motionEstimationMCTF at 2160p, 8 neighbour pictures, one GPU: the device-controlled search (vvb_mctf_estimate_pyramid_dev, planes resident) against the round-1
host replay (mctf_host.estimate_pyramid over the same kernels) and the reference's own motionEstimationLuma on the host cores (oracle/_ref, bounded sample).
usage: python tools/mctf_bench.py [W H] [refs]"""
import ctypes, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))


def main():
    import torch
    import vvenc_b200 as V
    import vvenc_b200._lib as L
    from vvenc_b200 import mctf_host as MH
    from _mctf_multigpu_run import pictures
    from _libs import have_ref, refshim, P
    W = int(sys.argv[1]) if len(sys.argv) > 2 else 3840
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 2160
    nrefs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    unit = 8 if min(W, H) < 720 else 16
    add_level = W >= 1920
    org, refs = pictures(W, H, nrefs)
    eng = V.CostEngine(0)
    pad = 128
    eng.upload_plane(0, MH.pad_edge(org, pad), W, H, pad)
    for i, r in enumerate(refs):
        eng.upload_plane(1 + i, MH.pad_edge(r, pad), W, H, pad)
    ext = torch.cuda.ExternalStream(eng.stream, device=torch.device('cuda', 0))
    fh, fw = (H + unit - 1) // unit, (W + unit - 1) // unit
    dfield = torch.zeros((nrefs, fh * fw * 4), dtype=torch.int32, device='cuda')
    par = L.vvb_mctf_pyr_par(unit, int(add_level), 0, 0)

    def run_all():
        for i in range(nrefs):
            eng._chk(eng.lib.vvb_mctf_estimate_pyramid_dev(eng.h, 0, 1 + i, ctypes.byref(par), ctypes.c_void_p(dfield[i].data_ptr())))
    l0 = eng.launches
    run_all(); eng.synchronize()
    launches = (eng.launches - l0) // nrefs
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 3
    with torch.cuda.stream(ext):
        e0.record(ext)
        for _ in range(reps):
            run_all()
        e1.record(ext)
    eng.synchronize(); torch.cuda.synchronize()
    dev_s = e0.elapsed_time(e1) / reps / 1e3
    fields_dev = dfield.cpu().numpy().reshape(nrefs, fh, fw, 4)
    out = {'picture': '%dx%d' % (W, H), 'refs': nrefs, 'unit': unit, 'levels': 5 if add_level else 4, 'device_control_s': dev_s, 'launches_per_neighbour_picture': int(launches),
           'block_refs_per_s': fh * fw * nrefs / dev_s}
    # the round-1 path: host replay over the same error kernels (includes level-picture uploads and the table downloads), first neighbour picture only
    class S:
        def __init__(s): s.eng = eng
        def make_provider(s, o, r):
            s.eng.upload_plane(30, MH.pad_edge(o, pad), o.shape[1], o.shape[0], pad); s.eng.upload_plane(31, MH.pad_edge(r, pad), r.shape[1], r.shape[0], pad)
            return MH.EngineProvider(s.eng, 30, 31)
    t0 = time.perf_counter()
    f = MH.estimate_pyramid(S().make_provider, org, refs[0], unit_size=unit, add_level=add_level)
    host_s = time.perf_counter() - t0
    out['host_replay_s_per_neighbour_picture'] = host_s
    out['device_over_host_replay'] = host_s / (dev_s / nrefs)
    eq = np.array_equal(fields_dev[0, ..., 0], f['x']) and np.array_equal(fields_dev[0, ..., 1], f['y']) and np.array_equal(fields_dev[0, ..., 2], f['error']) and \
        np.array_equal(fields_dev[0, ..., 3] & 0xffff, f['rmsme'].astype(np.int32))
    out['device_field_equals_host_replay'] = bool(eq)
    out['fractional_vectors'] = int(((fields_dev[..., 0] & 15) | (fields_dev[..., 1] & 15)).astype(bool).sum())
    if have_ref():
        R = refshim(); R.refshim_set_simd(b'AVX2')
        # the reference member on the host, one thread, a bounded sample: a 960x544 crop of the same pictures (same per-block work)
        cw, ch = 960, 544
        co = np.ascontiguousarray(org[:ch, :cw]); cr = np.ascontiguousarray(refs[0][:ch, :cw])
        exp = np.zeros(((ch + unit - 1) // unit, (cw + unit - 1) // unit, 4), dtype=np.int32)
        t0 = time.perf_counter()
        R.refshim_mctf_estimate_pyramid(1, P(co), P(cr), cw, ch, 10, unit, 0, 0, 0, P(exp))
        dt = time.perf_counter() - t0
        out['cpu_reference_1thread_s_per_neighbour_picture_scaled'] = dt * (W * H) / (cw * ch)
        out['cpu_sample'] = '960x544 crop, 4 levels, AVX2 members, 1 thread; scaled by area'
    print(json.dumps(out))


if __name__ == '__main__':
    main()
