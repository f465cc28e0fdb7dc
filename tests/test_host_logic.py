"""CPU-only: host-side candidate enumerators against the reference's point patterns (InterSearch.cpp:557-758)."""
import numpy as np
from vvenc_b200 import candidates as cand


def test_diamond_point_counts():
    # d=1 -> 4 (8 with corners); 2 <= d <= 8 -> 8 (4 axis at d, 4 diagonal at d/2); d > 8 -> 16 (InterSearch.cpp:574-705)
    assert len(cand.tz_diamond_points(1)) == 4 and len(cand.tz_diamond_points(1, True)) == 8
    for d in (2, 4, 8):
        pts = cand.tz_diamond_points(d)
        assert len(pts) == 8 and (0, -d) in pts and (-d // 2, -d // 2) in pts and (d // 2, d // 2) in pts
    for d in (16, 32, 64):
        pts = cand.tz_diamond_points(d)
        assert len(pts) == 16 and len(set(pts)) == 16
        assert all(abs(x) + abs(y) == d for x, y in pts)      # all on the L1 diamond of radius d


def test_pattern_sizes_and_order():
    p = cand.tz_diamond_pattern(64)
    assert len(p) == 1 + 4 + 3 * 8 + 3 * 16
    assert (p['dx'][0], p['dy'][0]) == (0, 0) and (p['dx'][1], p['dy'][1]) == (0, -1)
    pr = cand.tz_diamond_pattern(64, raster_step=5)
    assert len(pr) == len(p) + 26 * 26


def test_window_clip():
    assert cand.full_search_window(0, 0, 32, 3840, 2160, 0, 0, 16, 16, 80) == (-32, 32, -32, 32)
    l, r, t, b = cand.full_search_window(0, 0, 128, 3840, 2160, 3824, 0, 16, 16, 80)
    assert r == 80 and l == -128 and t == -80


def test_quad_order_grid_is_a_permutation_of_the_raster_grid():
    for (b, w, h) in ((8, 64, 48), (16, 3840, 2160), (64, 3840, 2160), (32, 96, 96)):
        xs, ys = cand.quad_order_grid(b, w, h)
        want = set((x, y) for y in range(0, h - b + 1, b) for x in range(0, w - b + 1, b))
        assert len(xs) == len(want) and set(zip(xs.tolist(), ys.tolist())) == want
        # leading entries come in proper quads
        assert (xs[1], ys[1]) == (xs[0] + b, ys[0]) and (xs[2], ys[2]) == (xs[0], ys[0] + b) and (xs[3], ys[3]) == (xs[0] + b, ys[0] + b)


def test_pyramid_lists_parent_child_relation():
    for (base, levels, w, h) in ((8, 4, 192, 152), (8, 4, 3840, 2160), (16, 2, 96, 80)):
        lists = cand.pyramid_lists(base, levels, w, h)
        for l in range(levels):
            s = base << l
            xs, ys = lists[l]
            want = set((x, y) for y in range(0, (h // s) * s, s) for x in range(0, (w // s) * s, s))
            assert set(zip(xs.tolist(), ys.tolist())) == want and len(xs) == len(want)      # every level tiles the picture exactly once
        for l in range(levels - 1):
            cx, cy = lists[l]; px, py = lists[l + 1]; s = base << l
            assert len(cx) >= 4 * len(px)
            for j in range(len(px)):
                kids = [(int(cx[4 * j + i]), int(cy[4 * j + i])) for i in range(4)]
                assert kids == [(px[j], py[j]), (px[j] + s, py[j]), (px[j], py[j] + s), (px[j] + s, py[j] + s)]
