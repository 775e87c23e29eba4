"""ctypes wrapper of oracle/_ref/libsdpl_ref.so -- the reference's OWN sources (src/ORBextractor.cc, src/Lineextractor.cc,
3rdparty/line_descriptor/src/*.cpp) compiled unmodified against the OpenCV stand-in in oracle/refshim/.  TEST
INFRASTRUCTURE ONLY (same rule as oracle.py): tests pin oracle/*.cpp against it; `bench.py --impl reference` may time it."""
import ctypes as C
import os
import subprocess
import numpy as np

from .oracle import KP_DTYPE, KL_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libsdpl_ref.so")
_lib = None


def build(reference="/root/reference"):
    """make -C oracle/refshim: needs the reference checkout; on a box without it the prebuilt .so is kept."""
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "refshim"), "REF=" + reference], stdout=subprocess.DEVNULL)
    return _SO


def available():
    return os.path.exists(_SO)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        build()
    L = C.CDLL(_SO, mode=os.RTLD_NOW)
    vp, ci, cf = C.c_void_p, C.c_int, C.c_float
    L.ref_orb_create.restype = vp
    L.ref_orb_create.argtypes = [ci, cf, ci, ci, ci]
    L.ref_orb_destroy.argtypes = [vp]
    L.ref_orb_tables.argtypes = [vp] * 7
    L.ref_orb_extract.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci]
    L.ref_orb_level_size.argtypes = [vp, ci, vp, vp]
    L.ref_orb_level_padded.argtypes = [vp, ci, vp]
    L.ref_orb_distribute.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp, ci]
    L.ref_line_create.restype = vp
    L.ref_line_create.argtypes = [ci, ci, cf, ci, cf, ci]
    L.ref_line_destroy.argtypes = [vp]
    L.ref_line_extract.argtypes = [vp, vp, ci, ci, ci, vp, vp, ci]
    L.ref_line_tables.argtypes = [vp, vp, ci, ci, ci, vp, vp, vp, vp]
    L.ref_lbd_compute.argtypes = [vp, ci, ci, ci, vp, ci, vp]
    L.ref_hamming256.argtypes = [vp, vp]
    L.ref_matcher_knn.argtypes = [vp, ci, vp, ci, ci, vp, vp, vp]
    L.ref_matcher_match.argtypes = [vp, ci, vp, ci, vp, vp]
    L.ref_matcher_stored_knn.argtypes = [vp, ci, vp, ci, vp, ci, ci, vp, vp, vp]
    _lib = L
    return L


def _u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


class RefORBextractor:
    """SDPL_SLAM::ORBextractor (include/ORBextractor.h:33-99) as compiled from the reference source."""

    def __init__(self, nfeatures, scale, nlevels, ini_th, min_th):
        self.L = lib()
        self.nlevels, self.nfeatures = nlevels, nfeatures
        self.h = self.L.ref_orb_create(nfeatures, scale, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_orb_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        f = [np.zeros(n, np.float32) for _ in range(4)]
        quota, umax = np.zeros(n, np.int32), np.zeros(16, np.int32)
        self.L.ref_orb_tables(self.h, *[a.ctypes.data for a in f], quota.ctypes.data, umax.ctypes.data)
        return f[0], f[1], f[2], f[3], quota, umax

    def __call__(self, img):
        img = _u8(img)
        cap = self.nfeatures + 4 * self.nlevels + 64
        kps, desc = np.zeros(cap, KP_DTYPE), np.zeros((cap, 32), np.uint8)
        n = self.L.ref_orb_extract(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], kps.ctypes.data,
                                   desc.ctypes.data, cap)
        assert 0 <= n <= cap, n
        return kps[:n].copy(), desc[:n].copy()

    def level_padded(self, level):
        w, h = C.c_int(), C.c_int()
        self.L.ref_orb_level_size(self.h, level, C.byref(w), C.byref(h))
        out = np.zeros((h.value + 38, w.value + 38), np.uint8)
        self.L.ref_orb_level_padded(self.h, level, out.ctypes.data)
        return out

    def distribute(self, xs, ys, resp, min_x, max_x, min_y, max_y, n_features, level=0):
        xs, ys, resp = (np.ascontiguousarray(a, np.float32) for a in (xs, ys, resp))
        cap = len(xs) + 8
        ox, oy, orr = (np.zeros(cap, np.float32) for _ in range(3))
        n = self.L.ref_orb_distribute(self.h, xs.ctypes.data, ys.ctypes.data, resp.ctypes.data, len(xs), min_x, max_x, min_y,
                                      max_y, n_features, level, ox.ctypes.data, oy.ctypes.data, orr.ctypes.data, cap)
        return ox[:n], oy[:n], orr[:n]


class RefLineextractor:
    """SDPL_SLAM::Lineextractor (include/Lineextractor.h:51-87) as compiled from the reference source; its
    cv::LineSegmentDetector is oracle/lsd_oracle.cpp (OpenCV is not vendored by the reference)."""

    def __init__(self, nfeatures, refine, lsd_scale, nlevels, scale, extractor):
        self.L = lib()
        self.nlevels = nlevels
        self.h = self.L.ref_line_create(nfeatures, refine, lsd_scale, nlevels, scale, extractor)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_line_destroy(self.h)
            self.h = None

    def __call__(self, img, cap=8192):
        img = _u8(img)
        kls, desc = np.zeros(cap, KL_DTYPE), np.zeros((cap, 32), np.uint8)
        n = self.L.ref_line_extract(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], kls.ctypes.data,
                                    desc.ctypes.data, cap)
        assert 0 <= n <= cap, n
        return kls[:n].copy(), desc[:n].copy()

    def tables(self, img):
        img = _u8(img)
        f = [np.zeros(self.nlevels, np.float32) for _ in range(4)]
        self.L.ref_line_tables(self.h, img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], *[a.ctypes.data for a in f])
        return f


def lbd_compute(img, keylines):
    img = _u8(img)
    kls = np.ascontiguousarray(keylines, dtype=KL_DTYPE)
    desc = np.zeros((len(kls), 32), np.uint8)
    n = lib().ref_lbd_compute(img.ctypes.data, img.shape[1], img.shape[0], img.strides[0], kls.ctypes.data, len(kls),
                              desc.ctypes.data)
    return desc[:n]


def hamming256(a, b):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return lib().ref_hamming256(a.ctypes.data, b.ctypes.data)


def matcher_knn(q, t, k):
    """BinaryDescriptorMatcher::knnMatch(query, train, matches, k): (train[nq,k], distance[nq,k], counts[nq])."""
    q, t = np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(t, np.uint8)
    train, dist = np.zeros((len(q), k), np.int32), np.zeros((len(q), k), np.float32)
    counts = np.zeros(len(q), np.int32)
    lib().ref_matcher_knn(q.ctypes.data, len(q), t.ctypes.data, len(t), k, train.ctypes.data, dist.ctypes.data,
                          counts.ctypes.data)
    return train, dist, counts


def matcher_stored_knn(q, t1, t2, k):
    """add([t1, t2]); train(); knnMatch(query, matches, k) on the stored set: (trainIdx[nq,k], imgIdx[nq,k], distance[nq,k])."""
    q, t1, t2 = (np.ascontiguousarray(a, np.uint8) for a in (q, t1, t2))
    train, img, dist = np.zeros((len(q), k), np.int32), np.zeros((len(q), k), np.int32), np.zeros((len(q), k), np.float32)
    lib().ref_matcher_stored_knn(q.ctypes.data, len(q), t1.ctypes.data, len(t1), t2.ctypes.data, len(t2), k, train.ctypes.data,
                                 img.ctypes.data, dist.ctypes.data)
    return train, img, dist


class RefFrame:
    """The reference's own Frame constructor (src/Frame.cc:236-907, compiled unmodified) on one image and its mask / depth / flow planes:
    runs the reference's ORBextractor and Lineextractor and then the constructor's loops; exposes the public result vectors."""

    def __init__(self, gray, depth, flow, mask, orb=(2000, 1.2, 8, 20, 7), line=(0, 2, 0.8, 2, 2.0), th_depth=40.0, th_depth_obj=25.0,
                 use_sample_fea=0):
        from . import oracle as orc
        L = lib()
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.ref_frame_construct.restype = vp
        L.ref_frame_construct.argtypes = [vp, vp, vp, vp, ci, ci, ci, cf, ci, ci, ci, ci, ci, cf, ci, cf, cf, cf, ci]
        L.ref_frame_destroy.argtypes = [vp]
        L.ref_frame_counts.argtypes = [vp, vp]
        L.ref_frame_points.argtypes = [vp] * 6
        L.ref_frame_lines.argtypes = [vp] * 7
        L.ref_frame_objects.argtypes = [vp] * 6
        L.ref_frame_grid.argtypes = [vp] * 3
        L.ref_frame_features_in_area.argtypes = [vp, cf, cf, cf, ci, ci, vp, ci]
        self._L = L
        g = np.ascontiguousarray(gray, np.uint8); d = np.ascontiguousarray(depth, np.float32)
        f = np.ascontiguousarray(flow, np.float32); m = np.ascontiguousarray(mask, np.int32)
        h, w = g.shape
        self._keep = (g, d, f, m)
        self._h = L.ref_frame_construct(g.ctypes.data, d.ctypes.data, f.ctypes.data, m.ctypes.data, w, h, orb[0], orb[1], orb[2], orb[3], orb[4],
                                        line[0], line[1], line[2], line[3], line[4], th_depth, th_depth_obj, use_sample_fea)
        c = np.zeros(6, np.int32); L.ref_frame_counts(self._h, c.ctypes.data)
        n, nl, ns, nsl, no, _ = [int(v) for v in c]
        KP, KL = orc.KP_DTYPE, orc.KL_DTYPE
        z = lambda k, dt: np.zeros(max(k, 1), dt)
        keys, stat, cor, fn, sd = z(n, KP), z(ns, KP), z(ns, KP), np.zeros((max(ns, 1), 2), np.float32), z(ns, np.float32)
        L.ref_frame_points(self._h, keys.ctypes.data, stat.ctypes.data, cor.ctypes.data, fn.ctypes.data, sd.ctypes.data)
        self.mvKeys, self.mvStatKeysTmp, self.mvCorres, self.mvFlowNext, self.mvStatDepthTmp = keys[:n], stat[:ns], cor[:ns], fn[:ns], sd[:ns]
        fl, sl, cl = z(nl, KL), z(nsl, KL), z(nsl, KL)
        lfn, inf, lsd = np.zeros((max(nsl, 1), 4), np.float32), np.zeros((max(nsl, 1), 3), np.float64), np.zeros((max(nsl, 1), 2), np.float32)
        L.ref_frame_lines(self._h, fl.ctypes.data, sl.ctypes.data, cl.ctypes.data, lfn.ctypes.data, inf.ctypes.data, lsd.ctypes.data)
        self.mvKeys_Line, self.mvStatKeysLineTmp, self.mvCorresLine = fl[:nl], sl[:nsl], cl[:nsl]
        self.mvFlowNext_Line, self.mvInfiniteLinesCorr, self.mvStatDepthLineTmp = lfn[:nsl], inf[:nsl], lsd[:nsl]
        ok, oc, ofn, od, ol = z(no, KP), z(no, KP), np.zeros((max(no, 1), 2), np.float32), z(no, np.float32), z(no, np.int32)
        L.ref_frame_objects(self._h, ok.ctypes.data, oc.ctypes.data, ofn.ctypes.data, od.ctypes.data, ol.ctypes.data)
        self.mvObjKeys, self.mvObjCorres, self.mvObjFlowNext, self.mvObjDepth, self.vSemObjLabel = ok[:no], oc[:no], ofn[:no], od[:no], ol[:no]
        cs, items = np.zeros(64 * 48 + 1, np.int32), np.zeros(max(n, 1), np.int32)
        L.ref_frame_grid(self._h, cs.ctypes.data, items.ctypes.data)
        self.grid_cell_start, self.grid_items = cs, items[:cs[-1]]

    def GetFeaturesInArea(self, x, y, r, minLevel=-1, maxLevel=-1):
        out = np.zeros(max(len(self.mvKeys), 1), np.int32)
        n = self._L.ref_frame_features_in_area(self._h, float(x), float(y), float(r), int(minLevel), int(maxLevel), out.ctypes.data, len(out))
        return out[:n]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.ref_frame_destroy(h)
            self._h = None


def matcher_match(q, t):
    q, t = np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(t, np.uint8)
    train, dist = np.zeros(len(q), np.int32), np.zeros(len(q), np.float32)
    lib().ref_matcher_match(q.ctypes.data, len(q), t.ctypes.data, len(t), train.ctypes.data, dist.ctypes.data)
    return train, dist
