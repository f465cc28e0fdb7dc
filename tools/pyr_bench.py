#!/usr/bin/env python3
"""times vvb_sad_search_pyramid_dev alone on the bench geometry (3840x2160, 8/16/32/64, +-32): tuning aid for pyramid_kernels.cuh
usage: [VVB_PYR_THREADS=n] [VVB_PYRAMID=0|1] python tools/pyr_bench.py [reps]"""
import ctypes, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench as B
import vvenc_b200 as V

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
eng = V.CostEngine(0)
if os.environ.get('VVB_PYRAMID', '') != '':
    eng.set_pyramid_engine(int(os.environ['VVB_PYRAMID']))
lib = eng.lib
ext = torch.cuda.ExternalStream(eng.stream)
sets = []
for s in range(4):
    org, ref, S = B.synth_picture_pair(1234 + 17 * s)
    dorg = torch.from_numpy(org).cuda(); dref = torch.from_numpy(ref).cuda()
    base = (B.MARGIN * S + B.MARGIN) * 2
    eng.bind_plane_dev(2 * s, dorg.data_ptr() + base, S, B.W, B.H, B.MARGIN, 10); eng.bind_plane_dev(2 * s + 1, dref.data_ptr() + base, S, B.W, B.H, B.MARGIN, 10)
    sets.append((dorg, dref))
d_blocks, d_best, counts = [], [], []
for n in B.SIZES:
    xs, ys = B.block_grid(n)
    b = np.zeros(len(xs), dtype=V.BLOCK_DT)
    b['x'] = xs; b['y'] = ys; b['left'] = -32; b['right'] = 32; b['top'] = -32; b['bottom'] = 32
    d_blocks.append(torch.from_numpy(np.frombuffer(b.tobytes(), dtype=np.uint8).copy()).cuda()); d_best.append(torch.empty(len(b) * 16, dtype=torch.uint8, device='cuda')); counts.append(len(b))
pb = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in d_blocks]); po = (ctypes.c_void_p * 4)(*[t.data_ptr() for t in d_best]); cn = (ctypes.c_int * 4)(*counts)
me = eng.me_par(B.LAMBDA, 2, 0, 0, 1, 2)
def run(i):
    s = i % 4
    rc = lib.vvb_sad_search_pyramid_dev(eng.h, 2 * s, 2 * s + 1, 4, pb, cn, 8, ctypes.byref(me), 65, 65, po)
    assert rc == 0, lib.vvb_last_error(eng.h)
for i in range(4):
    run(i)
eng.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(ext):
    e0.record(ext)
    for i in range(reps):
        run(i)
    e1.record(ext)
eng.synchronize(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
chk = int(torch.cat(d_best).to(torch.int64).sum().item())
print(json.dumps({'threads': os.environ.get('VVB_PYR_THREADS', 'auto'), 'engine': os.environ.get('VVB_PYRAMID', '1'), 'ms': ms, 'Tpel_diff_s': 3840 * 2160 * 4225 / (ms * 1e-3) / 1e12, 'checksum': chk}))
