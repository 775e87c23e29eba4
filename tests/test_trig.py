"""include/sdpl_trig.h -- the strict-IEEE double sin / cos that the CUDA kernels and the oracle share (oracle decision ix):
accuracy against mpmath, agreement with the C library, quadrant / sign handling, and the generated table."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sincos(oracle, xs):
    L = oracle.lib()
    L.orc_sdpl_sincos.argtypes = [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    s, c = C.c_double(), C.c_double()
    out = np.empty((len(xs), 2))
    for i, x in enumerate(xs):
        L.orc_sdpl_sincos(float(x), C.byref(s), C.byref(c))
        out[i] = s.value, c.value
    return out


def test_accuracy_against_mpmath(oracle):
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(-1, 1, 1500), rng.uniform(-10, 10, 1500), rng.uniform(-2000, 2000, 500),
                         np.array([0.0, 1e-30, -1e-9, 0.126, 0.855469, 0.8554690001, np.pi / 2, np.pi, 3 * np.pi / 2, 2 * np.pi, 7.0, -7.0])])
    got = _sincos(oracle, xs)
    worst = 0.0
    for x, (s, c) in zip(xs, got):
        for v, f in ((s, mp.sin), (c, mp.cos)):
            t = f(mp.mpf(float(x)))
            ulp = np.spacing(abs(float(t))) if float(t) != 0 else 5e-324
            worst = max(worst, float(abs(mp.mpf(float(v)) - t) / mp.mpf(float(ulp))))
    assert worst < 0.6, worst


def test_agrees_with_libm_up_to_the_last_bit(oracle):
    rng = np.random.default_rng(2)
    xs = rng.uniform(-20, 20, 20000)
    got = _sincos(oracle, xs)
    ds = np.abs(got[:, 0] - np.sin(xs)) / np.spacing(np.abs(np.sin(xs)))
    dc = np.abs(got[:, 1] - np.cos(xs)) / np.spacing(np.abs(np.cos(xs)))
    assert ds.max() <= 1.0 and dc.max() <= 1.0
    assert (ds == 0).mean() > 0.99 and (dc == 0).mean() > 0.99


def test_table_is_what_the_generator_writes(tmp_path):
    """include/sdpl_trig_table.inc is generated (tools/gen_trig_table.py, mpmath): regenerating gives the same bytes."""
    src = open(os.path.join(ROOT, "include", "sdpl_trig_table.inc")).read()
    work = tmp_path / "include"
    work.mkdir()
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "gen_trig_table.py")], cwd=tmp_path, stdout=subprocess.DEVNULL)
    assert open(work / "sdpl_trig_table.inc").read() == src
