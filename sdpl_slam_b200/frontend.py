"""Host-side mirror of the reference front-end interface over the C ABI (include/sdpl_frontend.h).

Class and method names follow the reference (argyrissm/SDPL-SLAM):
  ORBextractor            include/ORBextractor.h:33-99    (operator() -> __call__)
  Lineextractor           include/Lineextractor.h:51-87   (operator() -> __call__)
  BinaryDescriptorMatcher 3rdparty/line_descriptor/include/line_descriptor/descriptor_custom.hpp:1015-1126
All compute happens in the sm_100a CUDA library `lib/libsdpl_frontend.so`; there is no CPU fallback and the
import of this module fails loudly when the library is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsdpl_frontend.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
KL_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("pt_x", "<f4"), ("pt_y", "<f4"),
                     ("response", "<f4"), ("size", "<f4"), ("sx", "<f4"), ("sy", "<f4"), ("ex", "<f4"), ("ey", "<f4"),
                     ("sx_oct", "<f4"), ("sy_oct", "<f4"), ("ex_oct", "<f4"), ("ey_oct", "<f4"), ("length", "<f4"),
                     ("num_pixels", "<i4")])
DM_DTYPE = np.dtype([("query", "<i4"), ("train", "<i4"), ("img", "<i4"), ("distance", "<f4")])

SDPL_OK, SDPL_ERR_ARG, SDPL_ERR_CUDA, SDPL_ERR_CAPACITY, SDPL_ERR_OVERFLOW, SDPL_ERR_UNSUPPORTED = range(6)


class SdplError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sdpl_frontend error %d (%s): %s" % (code, _strerror(code), msg))
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen the CUDA front-end.  Raises (never falls back) when the library is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError("sdpl_slam_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(sdpl_slam_b200/csrc/build.sh).  There is no CPU fallback." % p)
    L = C.CDLL(p)
    vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
    ip = C.POINTER(C.c_int)
    L.sdpl_strerror.restype = C.c_char_p; L.sdpl_strerror.argtypes = [i]
    L.sdpl_last_error.restype = C.c_char_p
    L.sdpl_device_count.restype = i
    # ORB
    L.sdpl_orb_create.argtypes = [C.POINTER(vp), i, f, i, i, i, i]
    L.sdpl_orb_destroy.argtypes = [vp]; L.sdpl_orb_destroy.restype = None
    L.sdpl_orb_levels.argtypes = [vp]
    L.sdpl_orb_tables.argtypes = [vp, vp, vp, vp, vp]
    L.sdpl_orb_quota.argtypes = [vp, vp, vp]
    L.sdpl_orb_max_keypoints.argtypes = [vp]
    L.sdpl_orb_extract.argtypes = [vp, vp, i, i, i, vp, vp, i, ip]
    L.sdpl_orb_extract_batch.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, vp]
    L.sdpl_orb_extract_batch_dev.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, vp, i]
    L.sdpl_orb_pyramid_level.argtypes = [vp, i, i, vp, i, ip, ip]
    L.sdpl_orb_blurred_level.argtypes = [vp, i, i, vp, i]
    L.sdpl_orb_candidates.argtypes = [vp, i, i, vp, vp, vp, i, ip]
    L.sdpl_orb_level_counts.argtypes = [vp, i, vp]
    L.sdpl_orb_last_launches.argtypes = [vp]
    L.sdpl_orb_set_stream.argtypes = [vp, vp]
    for name in ("orb", "line", "matcher"):
        getattr(L, "sdpl_%s_set_profiling" % name).argtypes = [vp, i]
        getattr(L, "sdpl_%s_stage_times" % name).argtypes = [vp, vp, vp, vp, i]
    # lines
    L.sdpl_line_create.argtypes = [C.POINTER(vp), i, i, f, i, f, i, i]
    L.sdpl_line_destroy.argtypes = [vp]; L.sdpl_line_destroy.restype = None
    L.sdpl_line_levels.argtypes = [vp]
    L.sdpl_line_tables.argtypes = [vp, vp, vp, vp, vp]
    L.sdpl_line_extract.argtypes = [vp, vp, i, i, i, vp, vp, i, ip]
    L.sdpl_line_extract_batch.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, vp]
    L.sdpl_line_extract_batch_dev.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, i, vp, i]
    L.sdpl_line_lbd_compute.argtypes = [vp, vp, i, i, i, vp, i, vp]
    L.sdpl_line_lsd_segments.argtypes = [vp, i, i, vp, i, ip]
    L.sdpl_line_last_launches.argtypes = [vp]
    L.sdpl_line_set_stream.argtypes = [vp, vp]
    L.sdpl_line_set_serial.argtypes = [vp, i]
    L.sdpl_line_set_extractor.argtypes = [vp, i]
    L.sdpl_line_debug_grow_profile.argtypes = [vp, i, i, vp]
    L.sdpl_line_debug_grow_detail.argtypes = [vp, i]
    L.sdpl_line_debug_pending.argtypes = [vp, i, i, vp, i, ip]
    # matcher
    L.sdpl_matcher_create.argtypes = [C.POINTER(vp), i]
    L.sdpl_matcher_destroy.argtypes = [vp]; L.sdpl_matcher_destroy.restype = None
    L.sdpl_match_knn2.argtypes = [vp, vp, i, vp, i, vp, vp]
    L.sdpl_match_ratio.argtypes = [vp, vp, i, vp, i, f, i, vp, ip]
    L.sdpl_match_radius.argtypes = [vp, vp, i, vp, i, i, i, vp, vp]
    L.sdpl_match_knn2_batch_dev.argtypes = [vp, vp, vp, sz, vp, vp, sz, i, i, i, vp, vp, i]
    L.sdpl_match_ratio_batch_dev.argtypes = [vp, vp, vp, vp, i, i, f, i, vp, vp, i]
    L.sdpl_matcher_last_launches.argtypes = [vp]
    L.sdpl_match_knn.argtypes = [vp, vp, i, vp, i, i, vp]
    L.sdpl_matcher_add.argtypes = [vp, vp, i]; L.sdpl_matcher_add_dev.argtypes = [vp, vp, i]
    L.sdpl_matcher_train.argtypes = [vp]; L.sdpl_matcher_clear.argtypes = [vp]
    L.sdpl_matcher_train_size.argtypes = [vp, ip, ip]
    L.sdpl_matcher_knn.argtypes = [vp, vp, i, i, vp]
    L.sdpl_matcher_radius.argtypes = [vp, vp, i, i, i, vp, vp]
    L.sdpl_matcher_train_dev.argtypes = [vp, C.POINTER(vp), ip]
    L.sdpl_matcher_set_stream.argtypes = [vp, vp]
    L.sdpl_orb_check.argtypes = [vp]; L.sdpl_line_check.argtypes = [vp]
    L.sdpl_frontend_create.argtypes = [C.POINTER(vp), i, f, i, i, i, i, i, f, i, f, f, i, i]
    L.sdpl_frontend_destroy.argtypes = [vp]; L.sdpl_frontend_destroy.restype = None
    L.sdpl_frontend_capacities.argtypes = [vp, ip, ip]
    L.sdpl_frontend_set_line_capacity.argtypes = [vp, i]
    L.sdpl_frontend_set_line_extractor.argtypes = [vp, i]
    L.sdpl_frontend_reset.argtypes = [vp]
    L.sdpl_frontend_process.argtypes = [vp, vp, i, i, i, i, sz, vp, vp, vp, vp, vp, vp, vp]
    L.sdpl_frontend_last_launches.argtypes = [vp]
    L.sdpl_frontend_submit.argtypes = [vp, vp, i, i, i, i, sz]
    L.sdpl_frontend_collect.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ip]
    L.sdpl_frontend_pending.argtypes = [vp]
    L.sdpl_post_create.argtypes = [C.POINTER(vp), i]
    L.sdpl_post_destroy.argtypes = [vp]; L.sdpl_post_destroy.restype = None
    L.sdpl_post_set_stream.argtypes = [vp, vp]; L.sdpl_post_last_launches.argtypes = [vp]
    L.sdpl_post_set_profiling.argtypes = [vp, i]; L.sdpl_post_stage_times.argtypes = [vp, vp, vp, vp, i]
    L.sdpl_post_sample_objects_dev.argtypes = [vp, vp, vp, vp, i, i, i, i, f, vp, vp, vp, vp, vp, i, vp, i]
    L.sdpl_post_sample_objects.argtypes = [vp, vp, vp, vp, i, i, i, f, vp, vp, vp, vp, vp, i, ip]
    L.sdpl_post_filter_lines_dev.argtypes = [vp, vp, vp, i, i, i, vp, vp, i, vp, vp, vp, i]
    L.sdpl_post_point_corres_dev.argtypes = [vp, vp, vp, vp, i, i, i, vp, vp, i, f, vp, vp, vp, vp, vp, vp, i]
    L.sdpl_post_line_corres_dev.argtypes = [vp, vp, vp, vp, i, i, i, vp, vp, i, f, vp, vp, vp, vp, vp, vp, vp, vp, vp, i]
    L.sdpl_post_grid_dev.argtypes = [vp, i, i, i, vp, vp, i, i, i, vp, vp, i]
    L.sdpl_post_features_in_area_dev.argtypes = [vp, i, i, i, vp, i, vp, vp, i, i, vp, i, vp, i, vp, i]
    L.sdpl_post_search_area_dev.argtypes = [vp, i, i, i, vp, vp, i, vp, vp, i, i, vp, vp, i, vp, i]
    L.sdpl_post_distinctive_descriptors_dev.argtypes = [vp, vp, vp, i, vp, vp, i]
    L.sdpl_post_predict_scale_dev.argtypes = [vp, vp, vp, i, f, i, vp, i]
    L.sdpl_rows_digest_dev.argtypes = [vp, i, sz, vp, i, i, C.c_ulonglong, vp, vp]
    if path is None:
        _lib = L
    return L


def _strerror(code):
    try:
        return load_library().sdpl_strerror(code).decode()
    except Exception:  # pragma: no cover
        return "?"


def _check(rc):
    if rc != SDPL_OK:
        raise SdplError(rc, load_library().sdpl_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class _Profiled:
    """set_profiling / stage_times shared by the three handle classes (sdpl_*_set_profiling, sdpl_*_stage_times)."""
    _kind = ""

    def set_profiling(self, on=True):
        _check(getattr(self._L, "sdpl_%s_set_profiling" % self._kind)(self._h, int(bool(on))))

    def stage_times(self, cap=64):
        """[(stage name, milliseconds, kernels launched)] of the last call (synchronises on its last event)."""
        ms = (C.c_float * cap)(); names = (C.c_char_p * cap)(); nl = (C.c_int * cap)()
        n = getattr(self._L, "sdpl_%s_stage_times" % self._kind)(self._h, ms, names, nl, cap)
        return [(names[k].decode(), float(ms[k]), int(nl[k])) for k in range(n)]


def _gray(image):
    img = np.asarray(image)
    if img.dtype != np.uint8 or img.ndim != 2:
        # the reference asserts CV_8UC1 (src/ORBextractor.cc:1042)
        raise TypeError("image must be a 2-D uint8 (CV_8UC1) array")
    if img.strides[1] != 1:
        img = np.ascontiguousarray(img)
    return img


class ORBextractor(_Profiled):
    _kind = "orb"
    """SDPL_SLAM::ORBextractor.  `extractor(image, mask) -> (keypoints, descriptors)`; keypoints is a structured
    array with cv::KeyPoint's fields, descriptors an (N, 32) uint8 array."""

    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        _check(self._L.sdpl_orb_create(C.byref(self._h), int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST),
                                       int(minThFAST), int(device)))
        self.nfeatures, self.nlevels, self.device = int(nfeatures), int(nlevels), int(device)
        self._scale = float(np.float32(scaleFactor))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.sdpl_orb_destroy(h)
            self._h = None

    # --- getters, include/ORBextractor.h:49-69 ---
    def GetLevels(self):
        return self._L.sdpl_orb_levels(self._h)

    def GetScaleFactor(self):
        return self._scale

    def _tables(self):
        n = self.nlevels
        a = [np.empty(n, np.float32) for _ in range(4)]
        _check(self._L.sdpl_orb_tables(self._h, *[_p(x) for x in a]))
        return a

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self):
        q = np.empty(self.nlevels, np.int32); u = np.empty(16, np.int32)
        _check(self._L.sdpl_orb_quota(self._h, _p(q), _p(u)))
        return q, u

    def max_keypoints(self):
        return self._L.sdpl_orb_max_keypoints(self._h)

    # --- operator() ---
    def __call__(self, image, mask=None):
        if image is None or np.asarray(image).size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)   # silent return, src/ORBextractor.cc:1038
        img = _gray(image)
        cap = self.max_keypoints()
        kps = np.empty(cap, KP_DTYPE); desc = np.empty((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(self._L.sdpl_orb_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(desc), cap,
                                        C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        """images: (B, H, W) uint8 host array -> list of (keypoints, descriptors)."""
        imgs = np.ascontiguousarray(images, dtype=np.uint8)
        assert imgs.ndim == 3
        B, H, W = imgs.shape
        cap = self.max_keypoints()
        kps = np.empty((B, cap), KP_DTYPE); desc = np.empty((B, cap, 32), np.uint8); n = np.zeros(B, np.int32)
        _check(self._L.sdpl_orb_extract_batch(self._h, _p(imgs), B, W, H, W, W * H, _p(kps), _p(desc), cap, _p(n)))
        return [(kps[f, :n[f]].copy(), desc[f, :n[f]].copy()) for f in range(B)]

    def extract_batch_dev(self, d_imgs, nframes, w, h, d_kps, d_desc, capacity, d_n, stride=None, frame_stride=None, sync=False):
        """Device-pointer throughput path (ints are raw device addresses, e.g. torch tensor.data_ptr())."""
        stride = w if stride is None else stride
        frame_stride = stride * h if frame_stride is None else frame_stride
        _check(self._L.sdpl_orb_extract_batch_dev(self._h, C.c_void_p(d_imgs), nframes, w, h, stride, frame_stride,
                                                  C.c_void_p(d_kps), C.c_void_p(d_desc), capacity, C.c_void_p(d_n), int(sync)))

    def set_stream(self, cuda_stream):
        _check(self._L.sdpl_orb_set_stream(self._h, C.c_void_p(cuda_stream)))

    # --- mvImagePyramid and stage introspection (parity tests) ---
    def pyramid_level(self, level, frame=0):
        w = C.c_int(); h = C.c_int()
        _check(self._L.sdpl_orb_pyramid_level(self._h, frame, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value + 38, w.value + 38), np.uint8)
        _check(self._L.sdpl_orb_pyramid_level(self._h, frame, level, _p(out), out.strides[0], C.byref(w), C.byref(h)))
        return out

    def blurred_level(self, level, frame=0):
        w = C.c_int(); h = C.c_int()
        _check(self._L.sdpl_orb_pyramid_level(self._h, frame, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        _check(self._L.sdpl_orb_blurred_level(self._h, frame, level, _p(out), out.strides[0]))
        return out

    def candidates(self, level, frame=0):
        n = C.c_int()
        _check(self._L.sdpl_orb_candidates(self._h, frame, level, None, None, None, 0, C.byref(n)))
        xs = np.empty(n.value, np.int32); ys = np.empty(n.value, np.int32); rs = np.empty(n.value, np.int32)
        if n.value:
            _check(self._L.sdpl_orb_candidates(self._h, frame, level, _p(xs), _p(ys), _p(rs), n.value, C.byref(n)))
        return xs, ys, rs

    def level_counts(self, frame=0):
        c = np.empty(self.nlevels, np.int32)
        _check(self._L.sdpl_orb_level_counts(self._h, frame, _p(c)))
        return c

    def last_launches(self):
        return self._L.sdpl_orb_last_launches(self._h)


class Lineextractor(_Profiled):
    _kind = "line"
    """SDPL_SLAM::Lineextractor.  `extractor(image, mask) -> (keylines, descriptors_line)`."""

    def __init__(self, lsd_nfeatures=0, lsd_refine=2, lsd_scale=0.8, nlevels=2, scale=2.0, extractor=0, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        _check(self._L.sdpl_line_create(C.byref(self._h), int(lsd_nfeatures), int(lsd_refine), float(lsd_scale), int(nlevels),
                                        float(scale), int(extractor), int(device)))
        self.nlevels_l = int(nlevels)
        self.device = int(device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.sdpl_line_destroy(h)
            self._h = None

    def _tables(self):
        a = [np.empty(self.nlevels_l, np.float32) for _ in range(4)]
        _check(self._L.sdpl_line_tables(self._h, *[_p(x) for x in a]))
        return a

    # public members of the reference class (include/Lineextractor.h:66-75), as fixed tables
    @property
    def mvScaleFactor_l(self):
        return self._tables()[0]

    @property
    def mvInvScaleFactor_l(self):
        return self._tables()[1]

    @property
    def mvLevelSigma2_l(self):
        return self._tables()[2]

    @property
    def mvInvLevelSigma2_l(self):
        return self._tables()[3]

    def __call__(self, image, mask=None, capacity=16384):
        img = _gray(image)
        kls = np.empty(capacity, KL_DTYPE); desc = np.empty((capacity, 32), np.uint8)
        n = C.c_int(0)
        _check(self._L.sdpl_line_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kls), _p(desc),
                                         capacity, C.byref(n)))
        return kls[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, capacity=16384):
        imgs = np.ascontiguousarray(images, dtype=np.uint8)
        B, H, W = imgs.shape
        kls = np.empty((B, capacity), KL_DTYPE); desc = np.empty((B, capacity, 32), np.uint8); n = np.zeros(B, np.int32)
        _check(self._L.sdpl_line_extract_batch(self._h, _p(imgs), B, W, H, W, W * H, _p(kls), _p(desc), capacity, _p(n)))
        return [(kls[f, :n[f]].copy(), desc[f, :n[f]].copy()) for f in range(B)]

    def extract_batch_dev(self, d_imgs, nframes, w, h, d_kls, d_desc, capacity, d_n, stride=None, frame_stride=None, sync=False):
        stride = w if stride is None else stride
        frame_stride = stride * h if frame_stride is None else frame_stride
        _check(self._L.sdpl_line_extract_batch_dev(self._h, C.c_void_p(d_imgs), nframes, w, h, stride, frame_stride,
                                                   C.c_void_p(d_kls), C.c_void_p(d_desc), capacity, C.c_void_p(d_n), int(sync)))

    def compute(self, image, keylines):
        """BinaryDescriptor::compute on caller-provided keylines."""
        img = _gray(image)
        kls = np.ascontiguousarray(keylines, dtype=KL_DTYPE)
        desc = np.empty((kls.shape[0], 32), np.uint8)
        _check(self._L.sdpl_line_lbd_compute(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kls), kls.shape[0],
                                             _p(desc)))
        return desc

    def pending_rects(self, octave, frame=0, capacity=8192):
        out = np.zeros((capacity, 8), np.float64); n = C.c_int()
        _check(self._L.sdpl_line_debug_pending(self._h, frame, octave, _p(out), capacity, C.byref(n)))
        return out[:n.value].copy()

    def grow_detail(self, on=True):
        """detailed region-growing counters (grow_profile's phaseA ... phaseA_wait entries); off by default"""
        _check(self._L.sdpl_line_debug_grow_detail(self._h, int(on)))

    def grow_profile(self, octave, frame=0):
        out = np.zeros(16, np.int64)
        _check(self._L.sdpl_line_debug_grow_profile(self._h, frame, octave, _p(out)))
        return dict(zip(("select", "speculate", "commit", "rerun", "waves", "reruns", "dead", "seeds", "phaseA", "phaseB_busy", "phaseB_wait", "grow",
                         "rect_fit", "refine_tau", "grow_steps", "phaseA_wait"), out.tolist()))

    def set_serial(self, on=True):
        """region-growing schedule: 0/False block-level speculative waves (default), 1/True one seed at a time, 3 single-warp waves;
        `on | ((warps | ctas_per_sm << 4) << 8)` pins the kernel variant of mode 0, bit 2 selects the round-1 schedule"""
        _check(self._L.sdpl_line_set_serial(self._h, int(on)))

    def lsd_segments(self, octave, frame=0, capacity=65536):
        out = np.empty((capacity, 4), np.float32); n = C.c_int()
        _check(self._L.sdpl_line_lsd_segments(self._h, frame, octave, _p(out), capacity, C.byref(n)))
        return out[:n.value].copy()

    def set_stream(self, cuda_stream):
        _check(self._L.sdpl_line_set_stream(self._h, C.c_void_p(cuda_stream)))

    def last_launches(self):
        return self._L.sdpl_line_last_launches(self._h)


class BinaryDescriptorMatcher(_Profiled):
    _kind = "matcher"
    """Brute-force 256-bit Hamming matcher with the surface of cv::line_descriptor::BinaryDescriptorMatcher
    (match / knnMatch / radiusMatch).  Results are structured arrays with cv::DMatch's fields."""

    def __init__(self, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        _check(self._L.sdpl_matcher_create(C.byref(self._h), int(device)))
        self.device = int(device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.sdpl_matcher_destroy(h)
            self._h = None

    @staticmethod
    def _desc(d):
        d = np.ascontiguousarray(d, dtype=np.uint8)
        if d.ndim != 2 or d.shape[1] != 32:
            raise TypeError("descriptors must be (N, 32) uint8")
        return d

    # --- train set kept on the matcher, on the device: BinaryDescriptorMatcher::add / train / clear
    #     (descriptor_custom.hpp:1015-1126, binary_descriptor_matcher.cpp:127-194) ---
    def add(self, descriptors):
        """Append train descriptors (one (N, 32) array or a list of them, one per image).  Queries against the stored set
        return trainIdx = row in the concatenation of everything added and imgIdx = position of the image, as the reference."""
        ds = descriptors if isinstance(descriptors, (list, tuple)) else [descriptors]
        for d in ds:
            d = self._desc(d)
            _check(self._L.sdpl_matcher_add(self._h, _p(d), d.shape[0]))

    def train(self):
        """The reference builds its multi-index hash tables here; brute force needs no index."""
        _check(self._L.sdpl_matcher_train(self._h))

    def clear(self):
        _check(self._L.sdpl_matcher_clear(self._h))

    def train_size(self):
        a, b = C.c_int(), C.c_int()
        _check(self._L.sdpl_matcher_train_size(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def knnMatch(self, queryDescriptors, trainDescriptors=None, k=2):
        """k nearest train descriptors per query, ascending (distance, index).  k == 2 -> (best, second) arrays;
        other k -> an (nq, k) array (missing neighbours have train = -1).  Without trainDescriptors the stored set is used."""
        q = self._desc(queryDescriptors)
        k = int(k)
        if trainDescriptors is None:
            if self.train_size()[1] == 0:
                raise ValueError("no train descriptors: call add() first or pass trainDescriptors")
            out = np.zeros((q.shape[0], k), DM_DTYPE)
            _check(self._L.sdpl_matcher_knn(self._h, _p(q), q.shape[0], k, _p(out)))
            return (out[:, 0].copy(), out[:, 1].copy()) if k == 2 else out
        t = self._desc(trainDescriptors)
        if k != 2:
            out = np.zeros((q.shape[0], k), DM_DTYPE)
            _check(self._L.sdpl_match_knn(self._h, _p(q), q.shape[0], _p(t), t.shape[0], k, _p(out)))
            return out
        best = np.zeros(q.shape[0], DM_DTYPE); second = np.zeros(q.shape[0], DM_DTYPE)
        _check(self._L.sdpl_match_knn2(self._h, _p(q), q.shape[0], _p(t), t.shape[0], _p(best), _p(second)))
        return best, second

    def match(self, queryDescriptors, trainDescriptors=None):
        return self.knnMatch(queryDescriptors, trainDescriptors, 2)[0]

    def ratioMatch(self, queryDescriptors, trainDescriptors, ratio=0.8, max_dist=100):
        q, t = self._desc(queryDescriptors), self._desc(trainDescriptors)
        out = np.zeros(q.shape[0], DM_DTYPE); n = C.c_int(0)
        _check(self._L.sdpl_match_ratio(self._h, _p(q), q.shape[0], _p(t), t.shape[0], float(ratio), int(max_dist), _p(out),
                                        C.byref(n)))
        return out, n.value

    def radiusMatch(self, queryDescriptors, trainDescriptors, maxDistance, k=8):
        q = self._desc(queryDescriptors)
        counts = np.zeros(q.shape[0], np.int32); out = np.zeros((q.shape[0], k), DM_DTYPE)
        if trainDescriptors is None:
            _check(self._L.sdpl_matcher_radius(self._h, _p(q), q.shape[0], int(maxDistance), int(k), _p(counts), _p(out)))
            return counts, out
        t = self._desc(trainDescriptors)
        _check(self._L.sdpl_match_radius(self._h, _p(q), q.shape[0], _p(t), t.shape[0], int(maxDistance), int(k), _p(counts),
                                         _p(out)))
        return counts, out

    def knn2_batch_dev(self, d_q, d_nq, q_stride, d_t, d_nt, t_stride, npairs, max_q, max_t, d_best, d_second, sync=False):
        _check(self._L.sdpl_match_knn2_batch_dev(self._h, C.c_void_p(d_q), C.c_void_p(d_nq), q_stride, C.c_void_p(d_t),
                                                 C.c_void_p(d_nt), t_stride, npairs, max_q, max_t, C.c_void_p(d_best),
                                                 C.c_void_p(d_second), int(sync)))

    def ratio_batch_dev(self, d_best, d_second, d_nq, npairs, max_q, ratio, max_dist, d_out, d_n_acc, sync=False):
        _check(self._L.sdpl_match_ratio_batch_dev(self._h, C.c_void_p(d_best), C.c_void_p(d_second), C.c_void_p(d_nq), npairs, max_q,
                                                  float(ratio), int(max_dist), C.c_void_p(d_out), C.c_void_p(d_n_acc), int(sync)))

    def set_stream(self, cuda_stream):
        _check(self._L.sdpl_matcher_set_stream(self._h, C.c_void_p(cuda_stream)))

    def last_launches(self):
        return self._L.sdpl_matcher_last_launches(self._h)


FS_DTYPE = np.dtype([("n_kp", "<i4"), ("n_lines", "<i4"), ("n_pt_matches", "<i4"), ("n_ln_matches", "<i4")])


class FrontEnd:
    """The per-frame front-end in one call (sdpl_frontend_*): what Frame::Frame does with its two extractors
    (src/Frame.cc:314,328) plus frame-to-frame Hamming association of ORB and LBD descriptors.  process(images) takes a
    (B, H, W) uint8 host array of consecutive frames and returns padded host arrays plus per-frame counts."""

    def __init__(self, nfeatures=2000, scaleFactor=1.2, nlevels=8, iniThFAST=20, minThFAST=7, lsd_nfeatures=0, lsd_refine=2,
                 lsd_scale=0.8, lsd_levels=2, lsd_pyr_scale=2.0, ratio=0.8, max_dist=64, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        _check(self._L.sdpl_frontend_create(C.byref(self._h), nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, lsd_nfeatures,
                                            lsd_refine, lsd_scale, lsd_levels, lsd_pyr_scale, ratio, max_dist, device))
        a, b = C.c_int(), C.c_int()
        _check(self._L.sdpl_frontend_capacities(self._h, C.byref(a), C.byref(b)))
        self.kp_capacity, self.kl_capacity = a.value, b.value
        self._bufs = [None, None]
        self._flip = 0
        self._inflight = []

    def set_line_extractor(self, extractor):
        """0 = LSD (default), 1 = EDLines (Lineextractor's `extractor`, include/Lineextractor.h:60); only while nothing is in flight"""
        _check(self._L.sdpl_frontend_set_line_extractor(self._h, int(extractor)))

    def set_line_capacity(self, kl_capacity):
        """rows per frame of the key-line outputs (default 2048 when lsd_nfeatures == 0); only while nothing is in flight"""
        _check(self._L.sdpl_frontend_set_line_capacity(self._h, int(kl_capacity)))
        self.kl_capacity = int(kl_capacity)
        self._bufs = [None, None]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.sdpl_frontend_destroy(h)
            self._h = None

    def reset(self):
        _check(self._L.sdpl_frontend_reset(self._h))

    def _outputs(self, B):
        """two output sets, used alternately (a collected result stays valid until the collect after the next one)"""
        self._flip ^= 1
        cur = self._bufs[self._flip]
        if cur is None or cur[0].shape[0] < B:
            # only the set about to be written is (re)allocated: the other one still holds the previous result
            KC, LC = self.kp_capacity, self.kl_capacity
            cur = (np.empty((B, KC), KP_DTYPE), np.empty((B, KC, 32), np.uint8), np.empty((B, LC), KL_DTYPE),
                   np.empty((B, LC, 32), np.uint8), np.empty((B, KC), DM_DTYPE), np.empty((B, LC), DM_DTYPE),
                   np.empty(B, FS_DTYPE))
            self._bufs[self._flip] = cur
        return cur

    def submit(self, images):
        """Enqueue a (B, H, W) uint8 batch (upload, ORB || lines, matching) without waiting.  The array must stay alive and
        unchanged until the matching collect().  At most two batches in flight."""
        imgs = np.asarray(images)
        if imgs.dtype != np.uint8 or imgs.ndim != 3:
            raise TypeError("images must be a (B, H, W) uint8 array")
        if not imgs.flags["C_CONTIGUOUS"]:
            imgs = np.ascontiguousarray(imgs)
        B, H, W = imgs.shape
        _check(self._L.sdpl_frontend_submit(self._h, _p(imgs), B, W, H, W, W * H))
        self._inflight.append(imgs)       # only after the C FIFO accepted the batch

    def collect(self):
        """Results of the oldest submitted batch: dict(kps, desc, kls, ldesc, pt_matches, ln_matches: padded (B, cap, ...)
        arrays; stats: per-frame counts)."""
        if not self._inflight:
            raise SdplError(SDPL_ERR_ARG, "FrontEnd.collect: nothing submitted")
        imgs = self._inflight.pop(0)
        B = imgs.shape[0]
        kps, desc, kls, ldesc, pm, lm, st = self._outputs(B)
        n = C.c_int(0)
        _check(self._L.sdpl_frontend_collect(self._h, _p(kps), _p(desc), _p(kls), _p(ldesc), _p(pm), _p(lm), _p(st), C.byref(n)))
        return dict(kps=kps, desc=desc, kls=kls, ldesc=ldesc, pt_matches=pm, ln_matches=lm, stats=st[:n.value])

    def process(self, images):
        """submit + collect of one batch."""
        self.submit(images)
        return self.collect()

    def last_launches(self):
        return self._L.sdpl_frontend_last_launches(self._h)


class FramePost(_Profiled):
    _kind = "post"
    """The loops Frame::Frame runs on the extractor outputs (src/Frame.cc:349-389, 482-604, 728-809, 910-925) on the device.
    The *_dev methods take raw device addresses (ints, e.g. torch tensor.data_ptr()) of batched planes / feature blocks and are
    asynchronous on the handle's stream unless sync=True; sample_objects takes one frame's numpy planes."""

    def __init__(self, device=0):
        self._L = load_library()
        self._h = C.c_void_p()
        _check(self._L.sdpl_post_create(C.byref(self._h), int(device)))
        self.device = int(device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.sdpl_post_destroy(h)
            self._h = None

    def set_stream(self, cuda_stream):
        _check(self._L.sdpl_post_set_stream(self._h, C.c_void_p(cuda_stream)))

    def last_launches(self):
        return self._L.sdpl_post_last_launches(self._h)

    def sample_objects(self, maskSEM, imDepth, imFlow, step=4, thDepthObj=25.0):
        """Semi-dense features on objects of one frame (src/Frame.cc:769-809) -> dict(keys = mvObjKeys, corres = mvObjCorres,
        flow_next = mvObjFlowNext, depth = mvObjDepth, label = vSemObjLabel)."""
        m = np.ascontiguousarray(maskSEM, np.int32); d = np.ascontiguousarray(imDepth, np.float32); fl = np.ascontiguousarray(imFlow, np.float32)
        h, w = m.shape
        if d.shape != (h, w) or fl.shape != (h, w, 2):
            raise TypeError("maskSEM (h, w) int32, imDepth (h, w) float32 and imFlow (h, w, 2) float32 must agree in size")
        cap = ((h + step - 1) // step) * ((w + step - 1) // step)
        keys, corres = np.empty(cap, KP_DTYPE), np.empty(cap, KP_DTYPE)
        fn, dep, lab = np.empty((cap, 2), np.float32), np.empty(cap, np.float32), np.empty(cap, np.int32)
        n = C.c_int(0)
        _check(self._L.sdpl_post_sample_objects(self._h, _p(m), _p(d), _p(fl), w, h, int(step), float(thDepthObj), _p(keys), _p(corres), _p(fn),
                                                _p(dep), _p(lab), cap, C.byref(n)))
        k = n.value
        return dict(keys=keys[:k].copy(), corres=corres[:k].copy(), flow_next=fn[:k].copy(), depth=dep[:k].copy(), label=lab[:k].copy())

    def sample_objects_dev(self, d_mask, d_depth, d_flow, nframes, w, h, step, th_depth_obj, d_keys, d_corres, d_flow_next, d_depth_out, d_label,
                           capacity, d_n, sync=False):
        v = C.c_void_p
        _check(self._L.sdpl_post_sample_objects_dev(self._h, v(d_mask), v(d_depth), v(d_flow), nframes, w, h, step, float(th_depth_obj), v(d_keys),
                                                    v(d_corres), v(d_flow_next), v(d_depth_out), v(d_label), capacity, v(d_n), int(sync)))

    def filter_lines_dev(self, d_mask, d_depth, nframes, w, h, d_kls, d_n_in, capacity, d_out, d_keep_idx, d_n_out, sync=False):
        v = C.c_void_p
        _check(self._L.sdpl_post_filter_lines_dev(self._h, v(d_mask), v(d_depth), nframes, w, h, v(d_kls), v(d_n_in), capacity, v(d_out),
                                                  v(d_keep_idx), v(d_n_out), int(sync)))

    def point_corres_dev(self, d_mask, d_depth, d_flow, nframes, w, h, d_kps, d_n_in, capacity, th_depth, d_stat, d_corres, d_flow_next,
                         d_stat_depth, d_src_idx, d_n_out, sync=False):
        v = C.c_void_p
        _check(self._L.sdpl_post_point_corres_dev(self._h, v(d_mask), v(d_depth), v(d_flow), nframes, w, h, v(d_kps), v(d_n_in), capacity,
                                                  float(th_depth), v(d_stat), v(d_corres), v(d_flow_next), v(d_stat_depth), v(d_src_idx),
                                                  v(d_n_out), int(sync)))

    def line_corres_dev(self, d_mask, d_depth, d_flow, nframes, w, h, d_kls, d_n_in, capacity, th_depth, d_obj, d_n_obj, d_stat, d_corres,
                        d_flow_next, d_inf_line, d_stat_depth, d_src_idx, d_n_out, sync=False):
        v = C.c_void_p
        _check(self._L.sdpl_post_line_corres_dev(self._h, v(d_mask), v(d_depth), v(d_flow), nframes, w, h, v(d_kls), v(d_n_in), capacity,
                                                 float(th_depth), v(d_obj), v(d_n_obj), v(d_stat), v(d_corres), v(d_flow_next), v(d_inf_line),
                                                 v(d_stat_depth), v(d_src_idx), v(d_n_out), int(sync)))

    def grid_dev(self, nframes, w, h, d_kps, d_n_in, capacity, d_cell_start, d_items, grid_cols=64, grid_rows=48, sync=False):
        v = C.c_void_p
        _check(self._L.sdpl_post_grid_dev(self._h, nframes, w, h, v(d_kps), v(d_n_in), capacity, grid_cols, grid_rows, v(d_cell_start),
                                          v(d_items), int(sync)))


    def features_in_area_dev(self, nframes, w, h, d_kps, capacity, d_cell_start, d_items, d_queries, nq, d_out, max_out, d_counts, grid_cols=64,
                             grid_rows=48, sync=False):
        """Frame::GetFeaturesInArea for nq queries {x, y, r, minLevel, maxLevel} per frame (device pointers)."""
        v = C.c_void_p
        _check(self._L.sdpl_post_features_in_area_dev(self._h, nframes, w, h, v(d_kps), capacity, v(d_cell_start), v(d_items), grid_cols, grid_rows,
                                                      v(d_queries), nq, v(d_out), max_out, v(d_counts), int(sync)))


    def search_area_dev(self, nframes, w, h, d_kps, d_desc, capacity, d_cell_start, d_items, d_queries, d_qdesc, nq, d_out5, grid_cols=64, grid_rows=48,
                        sync=False):
        """best / second-best Hamming distance of nq query descriptors per frame inside their GetFeaturesInArea windows"""
        v = C.c_void_p
        _check(self._L.sdpl_post_search_area_dev(self._h, nframes, w, h, v(d_kps), v(d_desc), capacity, v(d_cell_start), v(d_items), grid_cols, grid_rows,
                                                 v(d_queries), v(d_qdesc), nq, v(d_out5), int(sync)))

    def distinctive_descriptors_dev(self, d_desc, d_start, n_points, d_best_idx, d_out_desc, sync=False):
        """MapPoint::ComputeDistinctiveDescriptors for n_points map points (CSR of observed descriptors)"""
        v = C.c_void_p
        _check(self._L.sdpl_post_distinctive_descriptors_dev(self._h, v(d_desc), v(d_start), n_points, v(d_best_idx), v(d_out_desc), int(sync)))

    def predict_scale_dev(self, d_max_distance, d_current_dist, n, log_scale_factor, n_levels, d_out, sync=False):
        v = C.c_void_p
        _check(self._L.sdpl_post_predict_scale_dev(self._h, v(d_max_distance), v(d_current_dist), n, float(log_scale_factor), n_levels, v(d_out),
                                                   int(sync)))


def rows_digest_dev(d_rows, row_bytes, frame_stride, d_n, nframes, max_rows, salt, d_digest, cuda_stream=0):
    """sdpl_rows_digest_dev: adds the 64-bit digest of every frame's valid rows to d_digest[f] (device pointers as ints)."""
    _check(load_library().sdpl_rows_digest_dev(C.c_void_p(d_rows), int(row_bytes), int(frame_stride), C.c_void_p(d_n), int(nframes),
                                               int(max_rows), int(salt), C.c_void_p(d_digest), C.c_void_p(cuda_stream)))


def device_count():
    return load_library().sdpl_device_count()
