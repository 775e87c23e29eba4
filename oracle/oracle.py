"""ctypes wrapper of the CPU ORACLE (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
KL_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
                     ("response", "<f4"), ("size", "<f4"), ("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"),
                     ("sx_oct", "<f4"), ("sy_oct", "<f4"), ("ex_oct", "<f4"), ("ey_oct", "<f4"), ("length", "<f4"),
                     ("num_pixels", "<i4")])
DM_DTYPE = np.dtype([("query", "<i4"), ("train", "<i4"), ("img", "<i4"), ("distance", "<f4")])
assert KP_DTYPE.itemsize == 28 and KL_DTYPE.itemsize == 68 and DM_DTYPE.itemsize == 16


def build(force=False, native=False, out=None):
    """Compile liboracle.so with g++ (make).  native=True builds a -march=native copy for CPU timing."""
    if native:
        out = out or os.path.join(_HERE, "_native", "liboracle.so")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        srcs = [os.path.join(_HERE, f) for f in sorted(os.listdir(_HERE)) if f.endswith(".cpp")]
        cmd = ["g++", "-O3", "-fPIC", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-march=native",
               "-shared", "-o", out] + srcs + ["-lm"]
        subprocess.check_call(cmd)
        return out
    args = ["make", "-C", _HERE] + (["-B"] if force else [])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, "liboracle.so")


_lib = None


def lib(path=None):
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(p):
        build()
    L = C.CDLL(p)
    u8p, i32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_float)
    L.orc_orb_create.restype = C.c_void_p
    L.orc_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
    L.orc_orb_destroy.argtypes = [C.c_void_p]
    L.orc_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
    L.orc_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_orb_level_size.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
    L.orc_orb_level_padded.restype = C.c_void_p
    L.orc_orb_level_padded.argtypes = [C.c_void_p, C.c_int]
    L.orc_orb_level_blurred.restype = C.c_void_p
    L.orc_orb_level_blurred.argtypes = [C.c_void_p, C.c_int]
    L.orc_orb_level_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_orb_level_count.argtypes = [C.c_void_p, C.c_int]
    L.orc_fast_atan2.restype = C.c_float
    L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
    L.orc_cv_round_f.argtypes = [C.c_float]
    L.orc_cv_round_d.argtypes = [C.c_double]
    L.orc_hamming256.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_match_ratio.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    if hasattr(L, "orc_line_create"):
        L.orc_line_create.restype = C.c_void_p
        L.orc_line_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_float, C.c_int]
        L.orc_line_destroy.argtypes = [C.c_void_p]
        L.orc_line_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_line_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.orc_line_last_segments.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    if hasattr(L, "orc_lsd_detect"):
        L.orc_lsd_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_lsd_last_scaled.argtypes = [C.c_void_p, C.c_int, i32p, i32p]
    if hasattr(L, "orc_lbd_compute"):
        L.orc_lbd_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    if path is None:
        _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


# ---------------- primitives ----------------
def resize_linear(img, dw, dh):
    img = _u8(img); out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw)
    return out


def resize_linear_exact(img, dw, dh, fx=0.0, fy=0.0):
    img = _u8(img); out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_exact_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw,
                                     C.c_double(fx), C.c_double(fy))
    return out


def border_reflect101(img, b):
    img = _u8(img); h, w = img.shape
    out = np.empty((h + 2 * b, w + 2 * b), np.uint8)
    lib().orc_border_reflect101_u8(_p(img), w, h, img.strides[0], _p(out), b, w + 2 * b)
    return out


def gaussian_blur(img, kind):
    img = _u8(img); out = np.empty_like(img)
    lib().orc_gaussian_blur_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), img.shape[1], kind)
    return out


def pyrdown(img, dw, dh):
    img = _u8(img); out = np.empty((dh, dw), np.uint8)
    lib().orc_pyrdown_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw)
    return out


def sobel3(img):
    img = _u8(img); dx = np.empty(img.shape, np.int16); dy = np.empty(img.shape, np.int16)
    lib().orc_sobel3_s16(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(dx), _p(dy))
    return dx, dy


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(float(y), float(x)))


def fast9_nms(img, th):
    img = _u8(img); cap = img.size
    xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32)
    n = lib().orc_fast9_nms(_p(img), img.shape[1], img.shape[0], img.strides[0], int(th), _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


# ---------------- ORB ----------------
class OrbOracle:
    """Mirror of SDPL_SLAM::ORBextractor (include/ORBextractor.h:33-99) on the oracle."""

    def __init__(self, nfeatures=2000, scale=1.2, nlevels=8, ini_th=20, min_th=7, _lib=None):
        self.L = _lib or lib()
        self.h = self.L.orc_orb_create(nfeatures, scale, nlevels, ini_th, min_th)
        assert self.h
        self.nfeatures, self.nlevels = nfeatures, nlevels

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_orb_destroy(self.h); self.h = None

    def tables(self):
        n = self.nlevels
        sf, isf, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        quota = np.empty(n, np.int32); umax = np.empty(16, np.int32)
        self.L.orc_orb_tables(self.h, _p(sf), _p(isf), _p(s2), _p(is2), _p(quota), _p(umax))
        return dict(scale=sf, inv_scale=isf, sigma2=s2, inv_sigma2=is2, quota=quota, umax=umax)

    def __call__(self, img, want_desc=True):
        img = _u8(img)
        cap = self.nfeatures + 4 * self.nlevels + 64
        kps = np.zeros(cap, KP_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps),
                                   _p(desc) if want_desc else None, cap)
        if n < 0:
            raise RuntimeError("oracle ORB failed: %d" % n)
        assert n <= cap
        return kps[:n].copy(), desc[:n].copy()

    def level_size(self, l):
        w = C.c_int32(); h = C.c_int32()
        self.L.orc_orb_level_size(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level_padded(self, l):
        w, h = self.level_size(l)
        ptr = self.L.orc_orb_level_padded(self.h, l)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(h + 38, w + 38)).copy()

    def level_blurred(self, l):
        w, h = self.level_size(l)
        ptr = self.L.orc_orb_level_blurred(self.h, l)
        if not ptr:
            return None
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(h, w)).copy()

    def level_candidates(self, l):
        cap = 1 << 20
        xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); rs = np.empty(cap, np.int32)
        n = self.L.orc_orb_level_candidates(self.h, l, _p(xs), _p(ys), _p(rs), cap)
        return xs[:n].copy(), ys[:n].copy(), rs[:n].copy()

    def level_count(self, l):
        return self.L.orc_orb_level_count(self.h, l)


# ---------------- lines ----------------
class LineOracle:
    """Mirror of SDPL_SLAM::Lineextractor (include/Lineextractor.h:51-87) on the oracle."""

    def __init__(self, nfeatures=0, refine=2, lsd_scale=0.8, nlevels=2, scale=2.0, extractor=0, _lib=None):
        self.L = _lib or lib()
        self.h = self.L.orc_line_create(nfeatures, refine, lsd_scale, nlevels, scale, extractor)
        assert self.h
        self.nlevels = nlevels

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_line_destroy(self.h); self.h = None

    def __call__(self, img, cap=20000):
        img = _u8(img)
        kls = np.zeros(cap, KL_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orc_line_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kls), _p(desc), cap)
        if n < 0:
            raise RuntimeError("oracle line extractor failed: %d" % n)
        assert n <= cap
        return kls[:n].copy(), desc[:n].copy()

    def last_segments(self, octave, cap=100000):
        """raw cv::LineSegmentDetector output (x1,y1,x2,y2) of pyramid octave `octave` of the last call"""
        out = np.empty((cap, 4), np.float32)
        n = self.L.orc_line_last_segments(self.h, int(octave), _p(out), cap)
        return out[:n].copy()

    def tables(self):
        n = self.nlevels
        a = [np.empty(n, np.float32) for _ in range(4)]
        self.L.orc_line_tables(self.h, *[_p(x) for x in a])
        return dict(scale=a[0], inv_scale=a[1], sigma2=a[2], inv_sigma2=a[3])


def lsd_detect(img, refine=2, scale=0.8, sigma_scale=0.6, quant=2.0, ang_th=22.5, log_eps=0.0, density_th=0.8,
               n_bins=1024, tie_mode=0, cap=100000):
    img = _u8(img)
    out = np.empty((cap, 4), np.float32)
    n = lib().orc_lsd_detect(_p(img), img.shape[1], img.shape[0], img.strides[0], refine, scale, sigma_scale, quant,
                             ang_th, log_eps, density_th, n_bins, tie_mode, _p(out), cap)
    assert 0 <= n <= cap
    return out[:n].copy()


def lbd_compute(img, kls, want_float=False):
    img = _u8(img); kls = np.ascontiguousarray(kls, dtype=KL_DTYPE)
    n = kls.shape[0]
    desc = np.zeros((n, 32), np.uint8); fd = np.zeros((n, 72), np.float32)
    lib().orc_lbd_compute(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(kls), n, _p(desc), _p(fd))
    return (desc, fd) if want_float else desc


# ---------------- matcher ----------------
def hamming256(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_hamming256(_p(a), _p(b))


def match_knn2(q, t):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    best = np.zeros(q.shape[0], DM_DTYPE); second = np.zeros(q.shape[0], DM_DTYPE)
    lib().orc_match_knn2(_p(q), q.shape[0], _p(t), t.shape[0], _p(best), _p(second))
    return best, second


def match_ratio(q, t, ratio=0.8, max_dist=100):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    out = np.zeros(q.shape[0], DM_DTYPE)
    n = lib().orc_match_ratio(_p(q), q.shape[0], _p(t), t.shape[0], float(ratio), int(max_dist), _p(out))
    return out, n


def match_stored_knn(q, train_sets, k):
    """BinaryDescriptorMatcher::add(train_sets) / train() / knnMatch(query, matches, k) on the stored set
    (binary_descriptor_matcher.cpp:127-194, 339-425): the k nearest rows of the CONCATENATION of all added images; trainIdx is
    the row of the concatenation (`results[j] - 1`, :393), imgIdx the image that row came from (`indexesMap.upper_bound(idx) - 1`,
    :381-383).  Returns an (nq, k) DMatch array; missing neighbours have train = -1."""
    sets = [np.ascontiguousarray(t, np.uint8).reshape(-1, 32) for t in train_sets]
    allt = np.concatenate(sets) if sets else np.zeros((0, 32), np.uint8)
    starts = np.cumsum([0] + [len(t) for t in sets])[:-1]
    _, out = match_radius(q, allt, 256, k)
    ok = out["train"] >= 0
    out["img"][ok] = np.searchsorted(starts, out["train"][ok], side="right") - 1
    return out


def match_radius(q, t, radius, k):
    q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
    counts = np.zeros(q.shape[0], np.int32); out = np.zeros((q.shape[0], k), DM_DTYPE)
    lib().orc_match_radius(_p(q), q.shape[0], _p(t), t.shape[0], int(radius), int(k), _p(counts), _p(out))
    return counts, out


# ---- Frame post-processing (oracle/post_oracle.cpp; reference: src/Frame.cc) ----
def _planes(mask, depth, flow):
    return (np.ascontiguousarray(mask, np.int32), np.ascontiguousarray(depth, np.float32), np.ascontiguousarray(flow, np.float32))


def post_sample_objects(mask, depth, flow, step=4, th_depth_obj=25.0):
    mask, depth, flow = _planes(mask, depth, flow)
    h, w = mask.shape
    cap = ((h + step - 1) // step) * ((w + step - 1) // step)
    keys, corres = np.zeros(cap, KP_DTYPE), np.zeros(cap, KP_DTYPE)
    fn, d, lab = np.zeros((cap, 2), np.float32), np.zeros(cap, np.float32), np.zeros(cap, np.int32)
    L = lib(); L.orc_post_sample_objects.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_float] + [C.c_void_p] * 5 + [C.c_int]
    n = L.orc_post_sample_objects(_p(mask), _p(depth), _p(flow), w, h, step, th_depth_obj, _p(keys), _p(corres), _p(fn), _p(d), _p(lab), cap)
    return dict(keys=keys[:n], corres=corres[:n], flow_next=fn[:n], depth=d[:n], label=lab[:n])


def post_filter_lines(kls, mask, depth):
    mask, depth, _ = _planes(mask, depth, np.zeros((1, 1, 2), np.float32))
    h, w = mask.shape
    kls = np.ascontiguousarray(kls, KL_DTYPE); n = len(kls)
    out, idx = np.zeros(max(n, 1), KL_DTYPE), np.zeros(max(n, 1), np.int32)
    L = lib(); L.orc_post_filter_lines.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    k = L.orc_post_filter_lines(_p(kls), n, _p(mask), _p(depth), w, h, _p(out), _p(idx))
    return out[:k], idx[:k]


def post_point_corres(kps, mask, depth, flow, th_depth=40.0):
    mask, depth, flow = _planes(mask, depth, flow)
    h, w = mask.shape
    kps = np.ascontiguousarray(kps, KP_DTYPE); n = len(kps); m = max(n, 1)
    stat, corres = np.zeros(m, KP_DTYPE), np.zeros(m, KP_DTYPE)
    fn, sd, idx = np.zeros((m, 2), np.float32), np.zeros(m, np.float32), np.zeros(m, np.int32)
    L = lib(); L.orc_post_point_corres.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_float] + [C.c_void_p] * 5
    k = L.orc_post_point_corres(_p(kps), n, _p(mask), _p(depth), _p(flow), w, h, th_depth, _p(stat), _p(corres), _p(fn), _p(sd), _p(idx))
    return dict(stat=stat[:k], corres=corres[:k], flow_next=fn[:k], depth=sd[:k], src_idx=idx[:k])


def post_line_corres(kls, mask, depth, flow, th_depth=40.0):
    mask, depth, flow = _planes(mask, depth, flow)
    h, w = mask.shape
    kls = np.ascontiguousarray(kls, KL_DTYPE); n = len(kls); m = max(n, 1)
    obj, stat, corres = np.zeros(m, KL_DTYPE), np.zeros(m, KL_DTYPE), np.zeros(m, KL_DTYPE)
    fn, inf, sd, idx = np.zeros((m, 4), np.float32), np.zeros((m, 3), np.float64), np.zeros((m, 2), np.float32), np.zeros(m, np.int32)
    nobj = C.c_int32(0)
    L = lib(); L.orc_post_line_corres.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_float] + [C.c_void_p] * 8
    k = L.orc_post_line_corres(_p(kls), n, _p(mask), _p(depth), _p(flow), w, h, th_depth, _p(obj), C.byref(nobj), _p(stat), _p(corres), _p(fn),
                               _p(inf), _p(sd), _p(idx))
    return dict(obj=obj[:nobj.value], stat=stat[:k], corres=corres[:k], flow_next=fn[:k], inf_line=inf[:k], depth=sd[:k], src_idx=idx[:k])


def post_grid(kps, w, h, grid_cols=64, grid_rows=48):
    kps = np.ascontiguousarray(kps, KP_DTYPE); n = len(kps)
    cs, items = np.zeros(grid_cols * grid_rows + 1, np.int32), np.zeros(max(n, 1), np.int32)
    L = lib(); L.orc_post_grid.argtypes = [C.c_void_p] + [C.c_int] * 5 + [C.c_void_p] * 2
    L.orc_post_grid(_p(kps), n, w, h, grid_cols, grid_rows, _p(cs), _p(items))
    return cs, items[:cs[-1]]


def post_features_in_area(kps, w, h, cell_start, items, x, y, r, min_level=-1, max_level=-1, grid_cols=64, grid_rows=48):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    cs = np.ascontiguousarray(cell_start, np.int32); it = np.ascontiguousarray(items, np.int32)
    out = np.zeros(max(len(kps), 1), np.int32)
    L = lib(); L.orc_post_features_in_area.argtypes = [C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_float] * 3 + [C.c_int] * 2 + [C.c_void_p, C.c_int]
    n = L.orc_post_features_in_area(_p(kps), w, h, grid_cols, grid_rows, _p(cs), _p(it), x, y, r, min_level, max_level, _p(out), len(out))
    return out[:n]


def post_distinctive_descriptors(desc, start):
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); start = np.ascontiguousarray(start, np.int32)
    n = len(start) - 1
    best, out = np.zeros(n, np.int32), np.zeros((n, 32), np.uint8)
    L = lib(); L.orc_post_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_post_distinctive_descriptors(_p(desc), _p(start), n, _p(best), _p(out))
    return best, out


def post_predict_scale(max_distance, current_dist, log_scale_factor, n_levels):
    a = np.ascontiguousarray(max_distance, np.float32); b = np.ascontiguousarray(current_dist, np.float32)
    out = np.zeros(len(a), np.int32)
    L = lib(); L.orc_post_predict_scale.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    L.orc_post_predict_scale(_p(a), _p(b), len(a), float(log_scale_factor), int(n_levels), _p(out))
    return out


def post_search_area(kps, desc, w, h, cell_start, items, x, y, r, min_level, max_level, qdesc, grid_cols=64, grid_rows=48):
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = np.ascontiguousarray(desc, np.uint8)
    cs = np.ascontiguousarray(cell_start, np.int32); it = np.ascontiguousarray(items, np.int32); q = np.ascontiguousarray(qdesc, np.uint8)
    out = np.zeros(5, np.int32)
    L = lib(); L.orc_post_search_area.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p] * 2 + [C.c_float] * 3 + [C.c_int] * 2 + [C.c_void_p] * 2
    L.orc_post_search_area(_p(kps), _p(desc), w, h, grid_cols, grid_rows, _p(cs), _p(it), x, y, r, min_level, max_level, _p(q), _p(out))
    return out
