"""Generates tests/golden/*.npz.  Run in the build container (needs python cv2 for the OpenCV-primitive vectors):
    python tools/make_golden.py
cv2_primitives.npz : outputs of the OpenCV primitives the reference calls, from cv2 itself, on seeded synth frames
                     (known-answer vectors that pin the oracle without cv2 at test time)
cv2_lsd.npz        : cv2.createLineSegmentDetector(2,0.8,0.6,2.0,22.5,0,0.8,1024).detect on seeded frames
frontend_small.npz : the oracle's end-to-end outputs (ORB keypoints/descriptors, keylines/LBD descriptors, matches) on
                     seeded frames -- regression vectors for the oracle and known answers for the CUDA path
Inputs are never stored: every image is synth.frame(seed, h, w), which is pure integer numpy (bit-reproducible)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cv2  # noqa: E402
from sdpl_slam_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

cv2.setNumThreads(1)
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def primitives():
    d = {}
    img = synth.frame(101, 96, 160, n_rect=6)
    d["meta_seed_h_w"] = np.array([101, 96, 160])
    d["resize_linear_133x80"] = cv2.resize(img, (133, 80), interpolation=cv2.INTER_LINEAR)
    d["resize_linear_80x48"] = cv2.resize(img, (80, 48), interpolation=cv2.INTER_LINEAR)     # exact 2x -> INTER_AREA path
    d["resize_exact_08"] = cv2.resize(img, None, fx=0.8, fy=0.8, interpolation=cv2.INTER_LINEAR_EXACT)
    d["border19"] = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    d["blur7_s2"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    d["blur5_s1"] = cv2.GaussianBlur(img, (5, 5), 1, 1, borderType=cv2.BORDER_REFLECT_101)
    d["blur7_s075"] = cv2.GaussianBlur(img, (7, 7), 0.75, 0.75, borderType=cv2.BORDER_REFLECT_101)
    d["pyrdown"] = cv2.pyrDown(img, dstsize=(80, 48))
    d["sobel_dx"] = cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=3)
    d["sobel_dy"] = cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=3)
    rng = np.random.default_rng(5)
    y = rng.integers(-3000, 3000, 512).astype(np.float32); x = rng.integers(-3000, 3000, 512).astype(np.float32)
    y[:8] = 0; x[4:12] = 0
    d["atan2_y"], d["atan2_x"] = y, x
    d["atan2_deg"] = np.array([cv2.fastAtan2(float(a), float(b)) for a, b in zip(y, x)], np.float32)
    for th in (20, 7):
        k = cv2.FastFeatureDetector_create(th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(img, None)
        d["fast%d" % th] = np.array([[int(p.pt[0]), int(p.pt[1]), int(p.response)] for p in k], np.int32).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, "cv2_primitives.npz"), **d)


def lsd():
    d = {}
    det = cv2.createLineSegmentDetector(2, 0.8, 0.6, 2.0, 22.5, 0.0, 0.8, 1024)
    cases = [(201, 120, 200), (202, 188, 621), (203, 240, 416)]
    d["cases"] = np.array(cases)
    for seed, h, w in cases:
        r = det.detect(synth.frame(seed, h, w))[0]
        d["lines_%d" % seed] = np.zeros((0, 4), np.float32) if r is None else r.reshape(-1, 4)
    np.savez_compressed(os.path.join(OUT, "cv2_lsd.npz"), **d)


def frontend():
    d = {}
    cases = [(301, 240, 416, 500, 8), (302, 188, 320, 300, 5)]
    d["cases"] = np.array(cases)
    for seed, h, w, nf, nl in cases:
        a, b = synth.frame(seed, h, w), synth.partner(seed, h, w)
        o = orc.OrbOracle(nf, 1.2, nl, 20, 7)
        ka, da = o(a); kb, db = o(b)
        kl, dl = orc.LineOracle(0, 2, 0.8, 2, 2.0, 0)(a)
        best, second = orc.match_knn2(da, db)
        d["kp_%d" % seed] = ka.view(np.uint8).reshape(-1, 28); d["desc_%d" % seed] = da
        d["kl_%d" % seed] = kl.view(np.uint8).reshape(-1, 68); d["ldesc_%d" % seed] = dl
        d["best_%d" % seed] = best.view(np.uint8).reshape(-1, 16); d["second_%d" % seed] = second.view(np.uint8).reshape(-1, 16)
    np.savez_compressed(os.path.join(OUT, "frontend_small.npz"), **d)


if __name__ == "__main__":
    primitives(); lsd(); frontend()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
