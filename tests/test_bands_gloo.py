"""CPU-only, world_size 2 over gloo: CTU-row band sharding + the single all-gather of per-row tables reproduce the
single-process result (the oracle stands in for the device kernels; the host-side N>1 logic is what is under test)."""
import os, sys, subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, 'tests')); sys.path.insert(0, %(root)r)
import cases as C
from _libs import oracle, P, PO
from vvenc_b200 import bands
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
sc = C.search_case(seed=99, W=128, H=96, margin=24)
S = sc['stride']; base = sc['margin'] * S + sc['margin']
y0, y1 = bands.split_ctu_rows(96, 32, 2)[rank]
xs, ys = bands.band_blocks(16, 128, y0, y1)
blk = np.zeros((len(xs), 10), dtype=np.int32); blk[:, 0] = xs; blk[:, 1] = ys; blk[:, 2] = 16; blk[:, 3] = 16
blk[:, 4] = -6; blk[:, 5] = 6; blk[:, 6] = -6; blk[:, 7] = 6
out = np.zeros((len(xs), 4), dtype=np.int32)
oracle().orc_full_search(PO(sc['org'], base), S, PO(sc['ref'], base), S, P(blk), len(xs), 0, 30.0, 2, 0, P(out), None, 0)
table = bands.all_gather_tables(torch.from_numpy(out.reshape(-1).copy()))
if rank == 0:
    np.save(sys.argv[2], table.numpy().reshape(-1, 4))
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_band_sharding_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import cases as C
    from _libs import oracle, P, PO
    from vvenc_b200 import bands
    port = 29600 + (os.getpid() % 200)
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT, 'port': port})
    outp = str(tmp_path / 'gathered.npy')
    procs = [subprocess.Popen([sys.executable, str(script), str(r), outp]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    got = np.load(outp)
    sc = C.search_case(seed=99, W=128, H=96, margin=24)
    S = sc['stride']; base = sc['margin'] * S + sc['margin']
    xs, ys = bands.band_blocks(16, 128, 0, 96)
    blk = np.zeros((len(xs), 10), dtype=np.int32); blk[:, 0] = xs; blk[:, 1] = ys; blk[:, 2] = 16; blk[:, 3] = 16
    blk[:, 4] = -6; blk[:, 5] = 6; blk[:, 6] = -6; blk[:, 7] = 6
    exp = np.zeros((len(xs), 4), dtype=np.int32)
    oracle().orc_full_search(PO(sc['org'], base), S, PO(sc['ref'], base), S, P(blk), len(xs), 0, 30.0, 2, 0, P(exp), None, 0)
    assert np.array_equal(got, exp)


def test_split_covers_picture():
    from vvenc_b200 import bands
    for (hgt, ctu, world) in ((2160, 128, 8), (2160, 64, 4), (4320, 128, 8), (240, 64, 8), (96, 32, 2)):
        b = bands.split_ctu_rows(hgt, ctu, world)
        assert b[0][0] == 0 and b[-1][1] == hgt and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert all((y0 % ctu == 0) for y0, _ in b if y0 < hgt)


MCTF_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(%(root)r, 'tests')); sys.path.insert(0, %(root)r)
from test_mctf_host import OracleProvider, mctf_shard_case
from vvenc_b200 import bands, mctf_host as MH
dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%(port)d', rank=int(sys.argv[1]), world_size=2)
org, refs, unit = mctf_shard_case()
H, W = org.shape
local = {}
for ref in bands.split_refs(len(refs), 2)[dist.get_rank()]:
    f = MH.estimate_pyramid(lambda o, r: OracleProvider(MH.pad_edge(o), MH.pad_edge(r), o.shape[1] + 256, 128), org, refs[ref], unit_size=unit)
    local[ref] = np.stack([f['x'], f['y'], f['error'], f['rmsme'].astype(np.int32)], axis=-1).reshape(-1, 4)
blocks = ((W + unit - 1) // unit) * ((H + unit - 1) // unit)
allf = bands.all_gather_motion_fields(local, len(refs), blocks)
if dist.get_rank() == 1:                       # any rank holds the complete set
    np.save(sys.argv[2], allf)
dist.barrier(); dist.destroy_process_group()
'''


def test_two_rank_mctf_reference_sharding_matches_single_process(tmp_path):
    """neighbour pictures dealt over the ranks, one all-gather of the motion fields; equals the single-process search of every picture"""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from test_mctf_host import OracleProvider, mctf_shard_case
    from vvenc_b200 import bands, mctf_host as MH
    port = 29850 + (os.getpid() % 100)
    script = tmp_path / 'mctf_worker.py'
    script.write_text(MCTF_WORKER % {'root': ROOT, 'port': port})
    outp = str(tmp_path / 'fields.npy')
    procs = [subprocess.Popen([sys.executable, str(script), str(r), outp]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    got = np.load(outp)
    org, refs, unit = mctf_shard_case()
    assert bands.split_refs(3, 2) == [[0, 2], [1]] and bands.split_refs(8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]
    for i, ref in enumerate(refs):
        f = MH.estimate_pyramid(lambda o, r: OracleProvider(MH.pad_edge(o), MH.pad_edge(r), o.shape[1] + 256, 128), org, ref, unit_size=unit)
        exp = np.stack([f['x'], f['y'], f['error'], f['rmsme'].astype(np.int32)], axis=-1).reshape(-1, 4)
        assert np.array_equal(got[i], exp), i
    assert np.any(got[:, :, :2] != 0)
