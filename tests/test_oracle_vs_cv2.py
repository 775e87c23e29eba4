"""Pins the CPU oracle at the OpenCV-primitive boundary: every primitive the reference calls in OpenCV (un-vendored;
README "tested with OpenCV 3.4") is compared bit-exactly with python cv2 on seeded images, and the LSD restatement is
compared with cv2.createLineSegmentDetector using the reference's parameters (LSDDetector_custom.cpp:291-298).
cv2 is only a test dependency; tests are skipped when it is not importable (the golden fixtures in tests/golden/
cover the same ground without cv2, see test_oracle_golden.py)."""
import numpy as np
import pytest

from sdpl_slam_b200 import synth

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module", autouse=True)
def _single_thread():
    cv2.setNumThreads(1)


def _img(seed, h=120, w=200):
    return synth.frame(seed, h, w, n_rect=8)


@pytest.mark.parametrize("shape,dst", [((375, 1242), (1035, 312)), ((120, 200), (167, 100)), ((375, 1242), (621, 188)),
                                       ((100, 200), (100, 50)), ((64, 64), (53, 53))])
def test_resize_linear(oracle, shape, dst):
    img = _img(1, *shape)
    np.testing.assert_array_equal(oracle.resize_linear(img, dst[0], dst[1]), cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("shape", [(375, 1242), (188, 621), (120, 200), (33, 47)])
def test_resize_linear_exact_08(oracle, shape):
    img = _img(2, *shape)
    ref = cv2.resize(img, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
    np.testing.assert_array_equal(oracle.resize_linear_exact(img, ref.shape[1], ref.shape[0], 0.8, 0.8), ref)


def test_border_reflect101(oracle):
    img = _img(3)
    np.testing.assert_array_equal(oracle.border_reflect101(img, 19), cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101))


@pytest.mark.parametrize("kind,ksize,sigma", [(0, 7, 2.0), (1, 5, 1.0), (2, 7, 0.75)])
def test_gaussian_blur(oracle, kind, ksize, sigma):
    for seed, shape in ((4, (120, 200)), (5, (375, 1242)), (6, (9, 11))):
        img = _img(seed, *shape)
        ref = cv2.GaussianBlur(img, (ksize, ksize), sigma, sigma, borderType=cv2.BORDER_REFLECT_101)
        np.testing.assert_array_equal(oracle.gaussian_blur(img, kind), ref)


def test_pyrdown_and_sobel(oracle):
    for seed, shape in ((7, (375, 1242)), (8, (121, 201))):
        img = _img(seed, *shape)
        h, w = img.shape
        np.testing.assert_array_equal(oracle.pyrdown(img, w // 2, h // 2), cv2.pyrDown(img, dstsize=(w // 2, h // 2)))
        dx, dy = oracle.sobel3(img)
        np.testing.assert_array_equal(dx, cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3))
        np.testing.assert_array_equal(dy, cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3))


def test_fast_atan2(oracle):
    rng = np.random.default_rng(9)
    y = rng.integers(-70000, 70000, 5000).astype(np.float32)
    x = rng.integers(-70000, 70000, 5000).astype(np.float32)
    y[:10] = 0; x[5:15] = 0
    ref = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)   # scalar cv::fastAtan2
    got = np.array([oracle.fast_atan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("th", [20, 7])
def test_fast9_nms(oracle, th):
    img = _img(10, 97, 131)
    fast = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = fast.detect(img, None)
    xs, ys, sc = oracle.fast9_nms(img, th)
    assert len(kps) == len(xs) > 0
    np.testing.assert_array_equal(xs, [int(k.pt[0]) for k in kps])
    np.testing.assert_array_equal(ys, [int(k.pt[1]) for k in kps])
    np.testing.assert_array_equal(sc, [int(k.response) for k in kps])


def test_cv_round_half_even(oracle):
    L = oracle.lib()
    assert [L.orc_cv_round_f(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5, 187.5)] == [0, 2, 2, 0, -2, 188]
    assert L.orc_cv_round_d(375 * 0.5) == 188


@pytest.mark.parametrize("seed,shape", [(1, (375, 1242)), (3, (480, 640)), (4, (188, 621)), (12, (97, 131))])
def test_lsd_matches_cv2(oracle, seed, shape):
    """cv::LineSegmentDetector with the reference's parameters (refine ADV, 0.8, 0.6, 2.0, 22.5, 0, 0.8, 1024)."""
    img = synth.frame(seed, *shape)
    lsd = cv2.createLineSegmentDetector(2, 0.8, 0.6, 2.0, 22.5, 0.0, 0.8, 1024)
    ref = lsd.detect(img)[0]
    ref = np.zeros((0, 4), np.float32) if ref is None else ref.reshape(-1, 4)
    got = oracle.lsd_detect(img)
    assert len(got) == len(ref)
    np.testing.assert_array_equal(got, ref)


def test_orb_keypoints_match_cv2_pipeline(oracle):
    """Level-0 cross-check of the ORB oracle's candidate stage against cv2.FAST run per 30-px cell as the reference does
    (src/ORBextractor.cc:778-818)."""
    img = synth.frame(13, 240, 416)
    o = oracle.OrbOracle(500, 1.2, 8, 20, 7)
    o(img)
    pad = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    np.testing.assert_array_equal(o.level_padded(0), pad)
    xs, ys, rs = o.level_candidates(0)
    h, w = img.shape
    minB, maxBX, maxBY = 16, w - 16, h - 16
    width, height = float(maxBX - minB), float(maxBY - minB)
    nCols, nRows = int(width / 30), int(height / 30)
    wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
    ex, ey, er = [], [], []
    f20 = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    f7 = cv2.FastFeatureDetector_create(7, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for i in range(nRows):
        iniY = minB + i * hCell; maxY = min(iniY + hCell + 6, maxBY)
        if iniY >= maxBY - 3:
            continue
        for j in range(nCols):
            iniX = minB + j * wCell; maxX = min(iniX + wCell + 6, maxBX)
            if iniX >= maxBX - 6:
                continue
            cell = np.ascontiguousarray(img[iniY:maxY, iniX:maxX])
            k = f20.detect(cell, None) or f7.detect(cell, None)
            for kp in k:
                ex.append(int(kp.pt[0]) + j * wCell); ey.append(int(kp.pt[1]) + i * hCell); er.append(int(kp.response))
    np.testing.assert_array_equal(xs, ex); np.testing.assert_array_equal(ys, ey); np.testing.assert_array_equal(rs, er)
