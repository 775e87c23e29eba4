"""GPU parity: the EDLines back-end of Lineextractor (extractor == 1: LSDDetectorC::detect_ED -> ED_Lib EDLines, then the shared key-line
construction and LBD) through the C ABI against the CPU oracle, which is pinned to the reference's own ED_Lib sources compiled
unmodified (tests/test_oracle_vs_ref.py).  Everything is integer or strict-IEEE double arithmetic on both sides, the transcendental
tables come from the host: key lines and LBD rows must be identical (angle: see test_gpu_line.py)."""
import numpy as np
import pytest

from sdpl_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _same(kg, dg, kr, dr):
    assert len(kg) == len(kr), "line count %d vs %d" % (len(kg), len(kr))
    for name in kr.dtype.names:
        if name == "angle":
            assert np.abs(kg[name] - kr[name]).max() <= 1e-6 if len(kg) else True
        else:
            np.testing.assert_array_equal(kg[name], kr[name], err_msg=name)
    np.testing.assert_array_equal(dg, dr)


@pytest.mark.parametrize("seed,h,w", [(1, 375, 1242), (2, 375, 1242), (3, 480, 640), (4, 240, 416), (5, 96, 160)])
def test_edlines_parity(frontend, oracle, seed, h, w):
    img = synth.frame(seed, h, w)
    kg, dg = frontend.Lineextractor(0, 2, 0.8, 2, 2.0, 1)(img)
    kr, dr = oracle.LineOracle(0, 2, 0.8, 2, 2.0, 1)(img)
    _same(kg, dg, kr, dr)
    assert len(kg) > 10 or h < 100


def test_edlines_batch_topn_and_octaves(frontend, oracle):
    imgs = synth.frames(range(70, 78), 375, 1242)
    gpu = frontend.Lineextractor(0, 2, 0.8, 2, 2.0, 1)
    ref = oracle.LineOracle(0, 2, 0.8, 2, 2.0, 1)
    res = gpu.extract_batch(imgs, capacity=4096)
    for f in range(len(imgs)):
        kr, dr = ref(imgs[f])
        _same(res[f][0], res[f][1], kr, dr)
    img = imgs[0]
    kg, dg = frontend.Lineextractor(40, 2, 0.8, 2, 2.0, 1)(img)
    kr, dr = oracle.LineOracle(40, 2, 0.8, 2, 2.0, 1)(img)
    assert len(kg) == 40
    _same(kg, dg, kr, dr)
    for nl in (1, 3):
        kg, dg = frontend.Lineextractor(0, 2, 0.8, nl, 2.0, 1)(img)
        kr, dr = oracle.LineOracle(0, 2, 0.8, nl, 2.0, 1)(img)
        _same(kg, dg, kr, dr)


def test_edlines_flat_image_and_highres(frontend, oracle):
    gpu = frontend.Lineextractor(0, 2, 0.8, 2, 2.0, 1)
    k, d = gpu(np.full((375, 1242), 80, np.uint8))
    assert len(k) == 0
    img = synth.frame(77, 1536, 2048)
    kg, dg = gpu(img, capacity=16384)
    kr, dr = oracle.LineOracle(0, 2, 0.8, 2, 2.0, 1)(img, cap=30000)
    _same(kg, dg, kr, dr)
    assert len(kg) > 100


def test_edlines_degenerate_joined_line(frontend, oracle):
    """Frame 2822 of the synthetic sequence: joining produces a line whose end points coincide on octave 1.  The reference divides by
    zero in EnumerateRectPoints and crashes there (decision 10: the line is not validated); oracle and kernel agree."""
    img = synth.sequence(2822, 2823, 375, 1242)[0]
    kg, dg = frontend.Lineextractor(0, 2, 0.8, 2, 2.0, 1)(img)
    kr, dr = oracle.LineOracle(0, 2, 0.8, 2, 2.0, 1)(img)
    _same(kg, dg, kr, dr)
    assert len(kg) == 293
