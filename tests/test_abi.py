"""CPU: the C-ABI library loads and exports every symbol include/sdpl_frontend.h declares; the host-side mirror
refuses to work without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sdpl_frontend.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdpl_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    from sdpl_slam_b200 import frontend as fe
    lib = ctypes.CDLL(fe.LIB_PATH)
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_pod_layouts_match_opencv_types():
    from sdpl_slam_b200 import frontend as fe
    assert fe.KP_DTYPE.itemsize == 28 and fe.KL_DTYPE.itemsize == 68 and fe.DM_DTYPE.itemsize == 16
    assert fe.KL_DTYPE.names[:3] == ("angle", "class_id", "octave") and fe.KL_DTYPE.names[-1] == "num_pixels"


def test_error_strings_and_no_cpu_fallback():
    from sdpl_slam_b200 import frontend as fe
    L = fe.load_library()
    assert L.sdpl_strerror(0) == b"ok" and b"no CPU fallback" in L.sdpl_strerror(2)
    if fe.device_count() > 0:
        pytest.skip("GPU present")
    for ctor in (lambda: fe.ORBextractor(2000, 1.2, 8, 20, 7), lambda: fe.BinaryDescriptorMatcher(),
                 lambda: fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0)):
        with pytest.raises(fe.SdplError) as ei:
            ctor()
        assert ei.value.code == fe.SDPL_ERR_CUDA


def test_missing_library_fails_loudly(tmp_path):
    from sdpl_slam_b200 import frontend as fe
    with pytest.raises(ImportError):
        fe.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sdpl_slam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_synth_is_deterministic():
    from sdpl_slam_b200 import synth
    a = synth.frame(3, 60, 80); b = synth.frame(3, 60, 80)
    assert (a == b).all() and a.dtype == np.uint8
    assert int(a.astype(np.int64).sum()) == int(synth.frame(3, 60, 80).astype(np.int64).sum())
    assert (synth.partner(3, 60, 80)[1:, 3:].astype(int) - a[:-1, :-3].astype(int) == 6).mean() > 0.9
