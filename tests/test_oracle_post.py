"""CPU: the Frame post-processing oracle (oracle/post_oracle.cpp, restating src/Frame.cc:349-389, 482-604, 728-809, 910-925) against
an independent numpy restatement of the same loops on synthetic planes: two restatements written separately must agree.  The pin
against the reference's own compiled Frame.cc is tests/test_oracle_vs_ref.py::test_frame_post_processing_equals_reference_frame_constructor;
this file also covers the functions restated from src/MapPoint.cc, which has no compiled form (not part of the reference's build)."""
import numpy as np

from sdpl_slam_b200 import synth

H, W = 188, 320


def _frame_features(oracle, seed):
    img = synth.frame(seed, H, W)
    kps, _ = oracle.OrbOracle(500, 1.2, 8, 20, 7)(img)
    kls, _ = oracle.LineOracle(0, 2, 0.8, 2, 2.0, 0)(img)
    return kps, kls


def test_object_sampling_equals_numpy(oracle):
    for seed in (1, 2):
        mask, depth, flow = synth.scene_planes(seed, H, W)
        r = oracle.post_sample_objects(mask, depth, flow, 4, 25.0)
        ii, jj = np.mgrid[0:H:4, 0:W:4]
        m, d, fx, fy = mask[ii, jj], depth[ii, jj], flow[ii, jj, 0], flow[ii, jj, 1]
        xj = jj.astype(np.float32) + fx; yi = ii.astype(np.float32) + fy
        keep = (m != 0) & (d < np.float32(25.0)) & (d > 0) & (xj < W) & (xj > 0) & (yi < H) & (yi > 0)
        assert keep.sum() > 50 and len(r["keys"]) == keep.sum()
        np.testing.assert_array_equal(r["keys"]["x"], jj[keep].astype(np.float32)); np.testing.assert_array_equal(r["keys"]["y"], ii[keep].astype(np.float32))
        np.testing.assert_array_equal(r["corres"]["x"], xj[keep]); np.testing.assert_array_equal(r["corres"]["y"], yi[keep])
        np.testing.assert_array_equal(r["flow_next"], np.stack([fx[keep], fy[keep]], 1))
        np.testing.assert_array_equal(r["depth"], d[keep]); np.testing.assert_array_equal(r["label"], m[keep])
        assert (r["keys"]["octave"] == -1).all() and (r["keys"]["class_id"] == -1).all() and (r["corres"]["size"] == 0).all()


def test_point_correspondences_equal_numpy(oracle):
    kps, _ = _frame_features(oracle, 5)
    mask, depth, flow = synth.scene_planes(5, H, W)
    r = oracle.post_point_corres(kps, mask, depth, flow, 40.0)
    x, y = kps["x"].astype(np.int32), kps["y"].astype(np.int32)          # truncation toward zero (coordinates are positive)
    d, fx, fy = depth[y, x], flow[y, x, 0], flow[y, x, 1]
    keep = (mask[y, x] == 0) & ~((d > np.float32(40.0)) | (d <= 0)) & (fx != 0) & (fy != 0) & (kps["x"] + fx < W) & (kps["y"] + fy < H)
    assert 20 < keep.sum() < len(kps)
    np.testing.assert_array_equal(r["src_idx"], np.nonzero(keep)[0])
    assert r["stat"].tobytes() == kps[keep].tobytes()
    np.testing.assert_array_equal(r["corres"]["x"], (kps["x"] + fx)[keep]); np.testing.assert_array_equal(r["corres"]["octave"], kps["octave"][keep])
    np.testing.assert_array_equal(r["depth"], d[keep])


def test_line_filters_and_correspondences_equal_numpy(oracle):
    _, kls = _frame_features(oracle, 6)
    mask, depth, flow = synth.scene_planes(6, H, W)
    out, idx = oracle.post_filter_lines(kls, mask, depth)
    x1, y1, x2, y2 = (kls[k].astype(np.int32) for k in ("sx", "sy", "ex", "ey"))
    xm, ym = (x1 + x2) // 2, (y1 + y2) // 2
    exp = (depth[y1, x1] + depth[y2, x2]) / np.float32(2)
    length = np.sqrt((x2 - x1).astype(np.float64) ** 2 + (y2 - y1).astype(np.float64) ** 2).astype(np.float32)
    thr = np.float32(10.0) * (length / np.float32(1000))
    keep = ~(np.abs(depth[ym, xm] - exp) > thr) & (mask[y1, x1] == mask[y2, x2])
    assert 5 < keep.sum() < len(kls)
    np.testing.assert_array_equal(idx, np.nonzero(keep)[0]); assert out.tobytes() == kls[keep].tobytes()
    r = oracle.post_line_corres(kls, mask, depth, flow, 40.0)
    ms, me = mask[y1, x1], mask[y2, x2]
    is_obj = (ms != 0) & (me != 0) & (ms == me)
    assert r["obj"].tobytes() == kls[is_obj].tobytes()
    ds, de = depth[y1, x1], depth[y2, x2]
    f = flow
    ok = (ms == 0) & (me == 0) & ~((x1 == x2) & (y1 == y2)) & ~((ds > 40) | (ds <= 0) | (de > 40) | (de <= 0))
    a, b, c, d_ = f[y1, x1, 0], f[y1, x1, 1], f[y2, x2, 0], f[y2, x2, 1]
    csx, csy, cex, cey = (u.astype(np.float32) + v for u, v in ((x1, a), (y1, b), (x2, c), (y2, d_)))      # int + float -> float
    ok &= (a != 0) & (b != 0) & (c != 0) & (d_ != 0) & (csx < W) & (csy < H) & (cex < W) & (cey < H) & (csx > 0) & (csy > 0) & (cex > 0) & (cey > 0)
    assert ok.sum() > 3
    np.testing.assert_array_equal(r["src_idx"], np.nonzero(ok)[0])
    np.testing.assert_array_equal(r["corres"]["sx"], csx[ok].astype(np.float32)); np.testing.assert_array_equal(r["corres"]["ey"], cey[ok].astype(np.float32))
    np.testing.assert_allclose(r["corres"]["angle"], np.arctan2((cey - csy)[ok].astype(np.float64), (cex - csx)[ok].astype(np.float64)), rtol=0, atol=5e-7)
    p = np.stack([csx[ok], csy[ok], np.ones(ok.sum(), np.float32)], 1).astype(np.float64)
    q = np.stack([cex[ok], cey[ok], np.ones(ok.sum(), np.float32)], 1).astype(np.float64)
    cr = np.cross(p, q); cr /= np.linalg.norm(cr, axis=1, keepdims=True)
    np.testing.assert_allclose(r["inf_line"], cr, rtol=1e-12, atol=1e-15)
    np.testing.assert_array_equal(r["depth"][:, 1], de[ok])


def test_grid_equals_numpy(oracle):
    kps, _ = _frame_features(oracle, 7)
    cs, items = oracle.post_grid(kps, W, H)
    px = np.round(kps["x"] * (np.float32(64) / np.float32(W))).astype(np.int32); py = np.round(kps["y"] * (np.float32(48) / np.float32(H))).astype(np.int32)
    # numpy rounds half to even, C round() half away from zero: the synthetic coordinates never sit on a half (checked)
    fracx = kps["x"] * (np.float32(64) / np.float32(W)); assert not np.any(np.abs(fracx - np.floor(fracx) - 0.5) < 1e-6)
    ok = (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
    cell = px * 48 + py
    want = [np.nonzero(ok & (cell == c))[0] for c in range(64 * 48)]
    assert cs[-1] == ok.sum()
    for c in (0, 17, 1000, 3071):
        np.testing.assert_array_equal(items[cs[c]:cs[c + 1]], want[c])
    np.testing.assert_array_equal(np.concatenate(want), items)


def test_features_in_area_equals_brute_force(oracle):
    """Frame::GetFeaturesInArea on the oracle's grid: the hits are exactly the key points inside the open square window with an
    admissible octave, ordered by (grid column, grid row, index) -- the order the reference's cell loops visit them."""
    kps, _ = _frame_features(oracle, 8)
    cs, items = oracle.post_grid(kps, W, H)
    px = np.round(kps["x"] * (np.float32(64) / np.float32(W))).astype(np.int32); py = np.round(kps["y"] * (np.float32(48) / np.float32(H))).astype(np.int32)
    rng = np.random.default_rng(3)
    for _ in range(40):
        x, y, r = np.float32(rng.uniform(-20, W + 20)), np.float32(rng.uniform(-20, H + 20)), np.float32(rng.uniform(1, 60))
        lo, hi = int(rng.integers(-1, 4)), int(rng.integers(-1, 8))
        got = oracle.post_features_in_area(kps, W, H, cs, items, float(x), float(y), float(r), lo, hi)
        inside = (np.abs(kps["x"] - x) < r) & (np.abs(kps["y"] - y) < r)
        if lo > 0 or hi >= 0:
            inside &= kps["octave"] >= lo
            if hi >= 0:
                inside &= kps["octave"] <= hi
        # a key point whose grid cell lies outside the cells the window covers is not found (the reference's cell arithmetic rounds
        # positions to the NEAREST cell but floors / ceils the window): reproduce the cell range
        wInv, hInv = np.float32(64) / np.float32(W), np.float32(48) / np.float32(H)
        x0, x1 = max(0, int(np.floor((x - r) * wInv))), min(63, int(np.ceil((x + r) * wInv)))
        y0, y1 = max(0, int(np.floor((y - r) * hInv))), min(47, int(np.ceil((y + r) * hInv)))
        inside &= (px >= x0) & (px <= x1) & (py >= y0) & (py <= y1) & (px >= 0) & (px < 64) & (py >= 0) & (py < 48)
        want = np.nonzero(inside)[0]
        want = want[np.lexsort((want, py[want], px[want]))]
        np.testing.assert_array_equal(got, want)


def test_distinctive_descriptor_and_predict_scale_equal_numpy(oracle):
    rng = np.random.default_rng(11)
    sizes = [1, 2, 3, 7, 20, 64, 0, 33]
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    base = rng.integers(0, 256, (len(sizes), 32), dtype=np.uint8)
    desc = np.zeros((start[-1], 32), np.uint8)
    for p, n in enumerate(sizes):
        for i in range(n):
            d = base[p].copy()
            for b in rng.choice(256, int(rng.integers(0, 40)), replace=False):
                d[b >> 3] ^= np.uint8(1 << (b & 7))
            desc[start[p] + i] = d
    best, out = oracle.post_distinctive_descriptors(desc, start)
    for p, n in enumerate(sizes):
        if n == 0:
            assert best[p] == -1; continue
        D = np.unpackbits(desc[start[p]:start[p + 1], None, :] ^ desc[None, start[p]:start[p + 1], :], axis=2).sum(2)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert best[p] == int(np.argmin(med)) and (out[p] == desc[start[p] + best[p]]).all()
    maxd = rng.uniform(1, 80, 5000).astype(np.float32); cur = rng.uniform(0.5, 90, 5000).astype(np.float32)
    got = oracle.post_predict_scale(maxd, cur, np.float32(np.log(np.float32(1.2))), 8)
    want = np.clip(np.ceil(np.log((maxd / cur).astype(np.float64)) / np.float64(np.float32(np.log(np.float32(1.2))))), 0, 7).astype(np.int32)
    assert (got != want).mean() < 1e-3 and np.abs(got - want).max() <= 1        # float vs double log: only on exact level boundaries
