"""seeded synthetic 8-bit 4:2:0 clips for the whole-encoder identity runs (the reference's own fixture test/data/RTn23_80x44p15_f15.yuv lives under /root/reference,
which the GPU box does not have; BASELINE configs[0]'s 416x240 clip is not in the reference tree at all)"""
import numpy as np


def write_clip(path, W, H, F, seed=1):
    rs = np.random.RandomState(seed)
    base = rs.randint(0, 256, size=(H // 4 + 8 + F, W // 4 + 8 + F)).astype(np.float32)
    up = np.kron(base, np.ones((4, 4), dtype=np.float32))
    sm = (up[:-4, :-4] + up[4:, :-4] + up[:-4, 4:] + up[4:, 4:] + 2 * up[2:-2, 2:-2]) / 6.0
    fr = []
    for f in range(F):
        y = np.clip(sm[8 + f:8 + f + H, 8 + 2 * f:8 + 2 * f + W] + rs.randint(-4, 5, size=(H, W)), 0, 255).astype(np.uint8)
        u = np.clip(sm[8:8 + H:2, 8 + 2 * f:8 + 2 * f + W:2] * 0.5 + 64, 0, 255).astype(np.uint8)
        v = np.clip(255 - sm[8:8 + H:2, 8 + 2 * f:8 + 2 * f + W:2] * 0.5, 0, 255).astype(np.uint8)
        fr.append(y.tobytes() + u.tobytes() + v.tobytes())
    with open(path, 'wb') as fh:
        fh.write(b''.join(fr))


def write_clip_scc(path, W, H, F, seed=1):
    """screen-content style clip (flat rectangles, thin strokes, isolated pixels, a global pan): makes transform skip win often enough to exercise its quantiser"""
    rs = np.random.RandomState(seed)
    base = np.full((H + 16, W + 16), 200, np.uint8)
    for _ in range(60):
        x, y = rs.randint(0, W), rs.randint(0, H); w, h = rs.randint(2, 24), rs.randint(1, 10)
        base[y:y + h, x:x + w] = rs.choice([0, 30, 90, 255, 140])
    for _ in range(400):
        x, y = rs.randint(0, W + 8), rs.randint(0, H + 8)
        base[y, x:x + rs.randint(1, 4)] = rs.choice([0, 255])
    fr = []
    for f in range(F):
        y = base[f:f + H, 2 * f:2 * f + W].copy()
        u = np.full((H // 2, W // 2), 128, np.uint8); u[::3, ::5] = 90
        v = np.full((H // 2, W // 2), 128, np.uint8); v[1::4, ::3] = 170
        fr.append(y.tobytes() + u.tobytes() + v.tobytes())
    with open(path, 'wb') as fh:
        fh.write(b''.join(fr))
