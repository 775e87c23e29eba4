"""CPU: the oracle against the committed golden vectors (tests/golden/, made by tools/make_golden.py from python cv2 --
the only executable form of the OpenCV primitives the reference calls -- and from the oracle itself for regression)."""
import os

import numpy as np
import pytest

from sdpl_slam_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(G, "cv2_primitives.npz"))


def test_primitives_known_answers(oracle, prim):
    seed, h, w = prim["meta_seed_h_w"]
    img = synth.frame(int(seed), int(h), int(w), n_rect=6)
    np.testing.assert_array_equal(oracle.resize_linear(img, 133, 80), prim["resize_linear_133x80"])
    np.testing.assert_array_equal(oracle.resize_linear(img, 80, 48), prim["resize_linear_80x48"])
    e = prim["resize_exact_08"]
    np.testing.assert_array_equal(oracle.resize_linear_exact(img, e.shape[1], e.shape[0], 0.8, 0.8), e)
    np.testing.assert_array_equal(oracle.border_reflect101(img, 19), prim["border19"])
    np.testing.assert_array_equal(oracle.gaussian_blur(img, 0), prim["blur7_s2"])
    np.testing.assert_array_equal(oracle.gaussian_blur(img, 1), prim["blur5_s1"])
    np.testing.assert_array_equal(oracle.gaussian_blur(img, 2), prim["blur7_s075"])
    np.testing.assert_array_equal(oracle.pyrdown(img, 80, 48), prim["pyrdown"])
    dx, dy = oracle.sobel3(img)
    np.testing.assert_array_equal(dx, prim["sobel_dx"]); np.testing.assert_array_equal(dy, prim["sobel_dy"])
    got = np.array([oracle.fast_atan2(float(a), float(b)) for a, b in zip(prim["atan2_y"], prim["atan2_x"])], np.float32)
    np.testing.assert_array_equal(got, prim["atan2_deg"])
    for th in (20, 7):
        xs, ys, sc = oracle.fast9_nms(img, th)
        np.testing.assert_array_equal(np.stack([xs, ys, sc], 1).reshape(-1, 3), prim["fast%d" % th])


def test_lsd_known_answers(oracle):
    g = np.load(os.path.join(G, "cv2_lsd.npz"))
    for seed, h, w in g["cases"]:
        got = oracle.lsd_detect(synth.frame(int(seed), int(h), int(w)))
        np.testing.assert_array_equal(got, g["lines_%d" % seed])


def test_frontend_regression(oracle):
    g = np.load(os.path.join(G, "frontend_small.npz"))
    for seed, h, w, nf, nl in g["cases"]:
        a, b = synth.frame(int(seed), int(h), int(w)), synth.partner(int(seed), int(h), int(w))
        o = oracle.OrbOracle(int(nf), 1.2, int(nl), 20, 7)
        ka, da = o(a); kb, db = o(b)
        np.testing.assert_array_equal(ka.view(np.uint8).reshape(-1, 28), g["kp_%d" % seed])
        np.testing.assert_array_equal(da, g["desc_%d" % seed])
        kl, dl = oracle.LineOracle(0, 2, 0.8, 2, 2.0, 0)(a)
        np.testing.assert_array_equal(kl.view(np.uint8).reshape(-1, 68), g["kl_%d" % seed])
        np.testing.assert_array_equal(dl, g["ldesc_%d" % seed])
        best, second = oracle.match_knn2(da, db)
        np.testing.assert_array_equal(best.view(np.uint8).reshape(-1, 16), g["best_%d" % seed])
        np.testing.assert_array_equal(second.view(np.uint8).reshape(-1, 16), g["second_%d" % seed])


def test_orb_tables_and_geometry(oracle):
    """SURVEY.md section 8 geometry block: level sizes, per-level quota, umax, pattern checksum."""
    import hashlib
    o = oracle.OrbOracle(2000, 1.2, 8, 20, 7)
    t = o.tables()
    np.testing.assert_array_equal(t["quota"], [434, 362, 302, 251, 209, 175, 145, 122])
    np.testing.assert_array_equal(t["umax"], [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3])
    o(synth.frame(1, 375, 1242))
    assert [o.level_size(l) for l in range(8)] == [(1242, 375), (1035, 312), (862, 260), (719, 217), (599, 181), (499, 151),
                                                    (416, 126), (347, 105)]
    np.testing.assert_array_equal(oracle.OrbOracle(2500, 1.2, 8, 20, 7).tables()["quota"], [543, 452, 377, 314, 262, 218, 182, 152])
    inc = open(os.path.join(os.path.dirname(G), "..", "include", "sdpl_orb_pattern.inc")).read()
    import re
    inc = re.sub(r"/\*.*?\*/", "", inc, flags=re.S)
    ints = [int(x) for x in re.findall(r"-?\d+", inc)]
    assert len(ints) == 1024
    assert hashlib.sha256(",".join(str(i) for i in ints).encode()).hexdigest() == \
        "88df8ca875cc8db56799edd57bb914edad8acb2d48c202b7a464a575b55dbdb8"


def test_matcher_oracle_properties(oracle):
    rng = np.random.default_rng(0)
    d = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    assert oracle.hamming256(d[0], d[0]) == 0
    assert oracle.hamming256(d[0], d[1]) == int(np.unpackbits(d[0] ^ d[1]).sum())
    b, s = oracle.match_knn2(d, d)
    assert (b["train"] == np.arange(300)).all() and (b["distance"] == 0).all()
    full = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(2)
    np.fill_diagonal(full, 1000)
    np.testing.assert_array_equal(s["distance"], full.min(1))
    np.testing.assert_array_equal(s["train"], full.argmin(1))       # ties -> lowest index
    b1, s1 = oracle.match_knn2(d[:5], d[:1])
    assert (s1["train"] == -1).all()


def test_line_oracle_edge_cases(oracle):
    flat = np.full((120, 200), 90, np.uint8)
    kl, dl = oracle.LineOracle()(flat)
    assert len(kl) == 0
    img = synth.frame(5, 120, 200)
    kl, dl = oracle.LineOracle()(img)
    assert len(kl) > 0 and (kl["class_id"] == np.arange(len(kl))).all()
    assert set(np.unique(kl["octave"])) <= {0, 1}
    assert (kl["length"] > 0.02 * 120).all()
    # top-N by response (Lineextractor.cc:73-82)
    kn, dn = oracle.LineOracle(nfeatures=5)(img)
    assert len(kn) == 5 and (np.diff(kn["response"]) <= 0).all() and (kn["class_id"] == np.arange(5)).all()
    # stand-alone LBD on the same keylines reproduces the descriptors
    np.testing.assert_array_equal(oracle.lbd_compute(img, kl), dl)
