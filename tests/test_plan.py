"""CPU tests of the call planner (csrc/engine.cu plan_call through rsr_debug_plan): the tile grid of the reference
(src/realsr.cpp:170-241), tiles grouped by padded shape, shape groups packed into jobs, the TTA slot layout, tile ranges for sharding.
Nothing here touches a GPU: the plan is what process() enqueues afterwards."""
import numpy as np
import pytest

import realsr_ncnn_vulkan_b200 as rb
from realsr_ncnn_vulkan_b200 import sharded

BYTES_PER_PX = 576 + 256 + 256 + 1024 + 4096  # workspace per network-input pixel (engine.cu plan_call)


def reference_tiles(w, h, T, P):
    """the reference's loop: for yi, for xi: tile origin, un-padded size, padded size (realsr.cpp:176-187, :235-241)"""
    out = []
    for yi in range((h + T - 1) // T):
        for xi in range((w + T - 1) // T):
            tw, th = min((xi + 1) * T, w) - xi * T, min((yi + 1) * T, h) - yi * T
            out.append((xi * T - P, yi * T - P, xi * T * 4, yi * T * 4, tw * 4, th * 4, th + 2 * P, tw + 2 * P))
    return out


def check_plan(e, w, h, first=0, count=-1):
    pl = e.plan(w, h, first, count)
    T, P = e.tilesize, e.prepadding
    ref = reference_tiles(w, h, T, P)
    want = ref[first:] if count < 0 else ref[first:first + count]
    d = pl["descs"]
    seen, px, jobs_bytes = [], 0, []
    for job in pl["jobs"]:
        jpx = 0
        for g in job["groups"]:
            assert g["px_off"] == jpx  # shape groups are concatenated pixel-major in every activation buffer
            jpx += g["n"] * g["H"] * g["W"]
            for t in d[g["off"]:g["off"] + g["n"]]:
                bh, bw = (t[8], t[7]) if t[6] >= 4 else (t[7], t[8])  # variants 4..7 run transposed
                assert (g["H"], g["W"]) == (bh, bw)
                seen.append(tuple(t))
        px += jpx
        jobs_bytes.append(jpx * (BYTES_PER_PX + (512 if e_tta(e) else 0)))
    assert px == pl["padded"]
    return pl, want, seen, jobs_bytes


def e_tta(e):
    return bool(getattr(e, "_tta", False))


def make(tta=False, tile=200, **opts):
    e = rb.RealSR(0, tta_mode=tta)
    e._tta = tta
    e.tilesize = tile
    for k, v in opts.items():
        assert e.set_option(k, v) == 0
    return e


@pytest.mark.parametrize("w,h,tile", [(1920, 1080, 400), (256, 256, 200), (3840, 2160, 200), (61, 45, 32), (1, 1, 200), (400, 400, 400), (401, 399, 400)])
def test_every_tile_runs_once_with_the_reference_geometry(w, h, tile):
    e = make(tile=tile)
    pl, want, seen, jobs_bytes = check_plan(e, w, h)
    assert len(seen) == len(want) == e.tile_count(w, h)
    got = sorted((t[0], t[1], t[2], t[3], t[4], t[5], t[7], t[8]) for t in seen)
    assert got == sorted(want)
    assert all(t[6] == 0 and t[9] == 0 for t in seen)
    assert pl["tta_bytes"] == 0 and not any(j["gathers"] for j in pl["jobs"])
    # the crops tile the output image exactly
    cover = np.zeros((4 * h, 4 * w), np.uint8)
    for t in seen:
        cover[t[3]:t[3] + t[5], t[2]:t[2] + t[4]] += 1
    assert (cover == 1).all()
    # merged launches: at most four shape groups per job, each shape at most once per job; 1080p at tile 400 is ONE job of four groups
    for j in pl["jobs"]:
        shapes = [(g["H"], g["W"]) for g in j["groups"]]
        assert 1 <= len(shapes) <= 4 and len(set(shapes)) == len(shapes)
    if (w, h, tile) == (1920, 1080, 400):
        assert len(pl["jobs"]) == 1 and sorted((g["H"], g["W"], g["n"]) for g in pl["jobs"][0]["groups"]) == [(300, 340, 1), (300, 420, 4), (420, 340, 2), (420, 420, 8)]
        assert pl["padded"] == 2302800
    assert all(b <= 16e9 for b in jobs_bytes) or len(seen) == 1


def test_job_packing_options():
    w, h = 3840, 2160  # 220 tiles of 220x220, 200x220 (bottom row: 2160 = 10*200 + 160 -> 180 high) ...
    base = check_plan(make(tile=200), w, h)
    one_shape = check_plan(make(tile=200, merge_shapes=0), w, h)
    assert all(len(j["groups"]) == 1 for j in one_shape[0]["jobs"])
    assert sorted(one_shape[2]) == sorted(base[2])  # same tiles, other packing
    small = check_plan(make(tile=200, max_batch=7), w, h)
    assert all(g["n"] <= 7 for j in small[0]["jobs"] for g in j["groups"]) and sorted(small[2]) == sorted(base[2])
    assert len(small[0]["jobs"]) > len(base[0]["jobs"])
    # the 16 GB workspace budget splits a 4K frame at tile 200 (10.09 M padded px x 6208 B = 62.6 GB) into several jobs
    assert len(base[0]["jobs"]) >= 4 and all(b <= 16e9 for b in base[3])


@pytest.mark.parametrize("w,h,tile", [(1280, 720, 200), (61, 45, 32), (100, 140, 200)])
def test_tta_plan_runs_eight_variants_per_tile_and_gathers_them(w, h, tile):
    e = make(tta=True, tile=tile)
    pl, want, seen, _ = check_plan(e, w, h)
    assert len(seen) == 8 * len(want) and pl["padded"] == 8 * sum(t[6] * t[7] for t in want)
    by_tile = {}
    for t in seen:
        by_tile.setdefault((t[0], t[1]), []).append(t)
    assert len(by_tile) == len(want)
    for ts in by_tile.values():
        assert sorted(t[6] for t in ts) == list(range(8))  # all eight dihedral variants (realsr.cpp:242-318)
    d = pl["descs"]
    for j in pl["jobs"]:
        assert len(j["groups"]) % 2 == 0  # variants 0..3 as H x W, 4..7 as W x H
        slots_needed = 0
        for ga in j["gathers"]:
            g8 = d[ga["off"]:ga["off"] + 8 * ga["n"]]
            assert list(g8[:, 6]) == list(range(8)) * ga["n"]                       # slot order: tile-major, variant-minor
            assert list(g8[:, 9]) == list(range(8 * ga["n"]))                       # tta_slot = index of the fp32 result plane
            assert ga["mw"] == g8[:, 4].max() and ga["mh"] == g8[:, 5].max()
            assert ga["slot_stride"] == 16 * g8[0, 7] * g8[0, 8]
            assert ga["tta_off"] == slots_needed
            slots_needed += 8 * ga["n"] * ga["slot_stride"]
        assert j["tta_float4"] == slots_needed and pl["tta_bytes"] >= 16 * slots_needed


def test_tile_ranges_of_the_sharded_call_partition_the_grid():
    """sharded.py cuts the tile grid into contiguous ranges of equal padded-pixel cost, one per GPU"""
    w, h, tile = 1920, 1080, 400
    e = make(tile=tile)
    n = e.tile_count(w, h)
    full = sorted(check_plan(e, w, h)[2])
    for parts in (1, 2, 3, 8, 15, 20):
        ranges = sharded.split_tiles(w, h, tile, e.prepadding, parts)
        assert ranges[0][0] == 0 and sum(c for _, c in ranges) == n
        assert all(ranges[i][0] + ranges[i][1] == ranges[i + 1][0] for i in range(len(ranges) - 1))
        seen, costs = [], []
        for first, count in ranges:
            if count == 0:
                continue
            pl, want, s, _ = check_plan(e, w, h, first, count)
            assert len(s) == count
            seen += s
            costs.append(pl["padded"])
        assert sorted(seen) == full
        if parts <= n:
            assert max(costs) <= 2.05 * (2302800 / parts) + 420 * 420  # balanced up to one tile


def test_plan_argument_checks():
    e = make(tile=200)
    with pytest.raises(rb.RealSRError):
        e.plan(0, 10)
    assert e.plan(50, 50, 5, 3)["descs"].shape == (0, 10)  # a range beyond the grid is empty, not an error
