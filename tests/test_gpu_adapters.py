"""include/sdpl_adapters.hpp -- the C++ classes with the reference's own signatures (SDPL_SLAM::ORBextractor,
SDPL_SLAM::Lineextractor, a BinaryDescriptorMatcher-shaped matcher) -- compiled against the reference's KeyLine header and
linked with the product library (oracle/refshim/adapter_check.cpp, built by oracle/refshim/Makefile into oracle/_ref/).
The binary calls them the way Frame::ExtractORB / ExtractLines do (src/Frame.cc:927-949); its output must be byte-identical
to what the C ABI returns through the Python mirror, and equal to the oracle."""
import os
import subprocess

import numpy as np
import pytest

from sdpl_slam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "adapter_check")


def _run(tmp_path, h, w, seed):
    a, b = synth.frame(seed, h, w), synth.partner(seed, h, w)
    pa, pb, po = tmp_path / "a.u8", tmp_path / "b.u8", tmp_path / "out.bin"
    a.tofile(pa); b.tofile(pb)
    out = subprocess.run([BIN, str(w), str(h), str(pa), str(pb), str(po)], capture_output=True, text=True, timeout=300)
    return a, b, po, out


def test_adapter_binary_fails_loudly_without_a_device(tmp_path):
    """On a box without a GPU the adapter classes must throw (no CPU fallback); on a GPU box this is covered below."""
    import torch
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/adapter_check not built (needs the reference checkout)")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _, _, _, out = _run(tmp_path, 120, 160, 1)
    assert out.returncode == 1 and "no CPU fallback" in out.stderr


@pytest.mark.gpu
def test_adapters_equal_c_abi_and_oracle(frontend, oracle, tmp_path):
    assert os.path.exists(BIN), "oracle/_ref/adapter_check missing: run __graft_entry__.build() where /root/reference exists"
    from sdpl_slam_b200.frontend import KP_DTYPE, KL_DTYPE, DM_DTYPE
    h, w = 375, 1242
    a, b, po, out = _run(tmp_path, h, w, 5)
    assert out.returncode == 0, out.stderr
    raw = po.read_bytes()
    nkp, nkl, nm = np.frombuffer(raw, np.int32, 3)
    off = 12
    kps = np.frombuffer(raw, KP_DTYPE, nkp, off); off += 28 * nkp
    desc = np.frombuffer(raw, np.uint8, nkp * 32, off).reshape(nkp, 32); off += 32 * nkp
    kls = np.frombuffer(raw, KL_DTYPE, nkl, off); off += 68 * nkl
    ldesc = np.frombuffer(raw, np.uint8, nkl * 32, off).reshape(nkl, 32); off += 32 * nkl
    best = np.frombuffer(raw, DM_DTYPE, nm, off); off += 16 * nm
    second = np.frombuffer(raw, DM_DTYPE, nm, off); off += 16 * nm
    assert off == len(raw) and nm == nkp
    orb = frontend.ORBextractor(2000, 1.2, 8, 20, 7)
    k1, d1 = orb(a)
    k2, d2 = orb(b)
    assert kps.tobytes() == k1.tobytes() and (desc == d1).all()
    gl, gd = frontend.Lineextractor(0, 2, 0.8, 2, 2.0, 0)(a)
    assert kls.tobytes() == gl.tobytes() and (ldesc == gd).all()
    gb, gs = frontend.BinaryDescriptorMatcher().knnMatch(d1, d2, 2)
    assert best.tobytes() == gb.tobytes() and second.tobytes() == gs.tobytes()
    # and against the oracle (ORB + matches bit-exact)
    ok, od = oracle.OrbOracle(2000, 1.2, 8, 20, 7)(a)
    for name in ("x", "y", "size", "response", "octave", "class_id"):
        assert (kps[name] == ok[name]).all(), name
    assert np.abs(kps["angle"] - ok["angle"]).max() <= 1e-3
    assert (desc == od).all()
    ok2, od2 = oracle.OrbOracle(2000, 1.2, 8, 20, 7)(b)
    rb, rs = oracle.match_knn2(od, od2)
    assert (best["train"] == rb["train"]).all() and (best["distance"] == rb["distance"]).all()
    assert (second["train"] == rs["train"]).all() and (second["distance"] == rs["distance"]).all()
