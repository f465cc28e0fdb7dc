#!/usr/bin/env python3
"""two device-controlled MCTF pyramids of one 2160p neighbour picture (for `ncu --metrics gpu__time_duration.sum`: the second one is the warm one)"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import vvenc_b200 as V
from vvenc_b200 import mctf_host as MH
from _mctf_multigpu_run import pictures
W, H = 3840, 2160
org, refs = pictures(W, H, 1)
eng = V.CostEngine(0)
eng.upload_plane(0, MH.pad_edge(org, 128), W, H, 128); eng.upload_plane(1, MH.pad_edge(refs[0], 128), W, H, 128)
for _ in range(2):
    f = eng.mctf_estimate_pyramid(0, 1, W, H, 16, True)
print(int((f['x'] != 0).sum()), int(((f['x'] | f['y']) & 15 != 0).sum()))
