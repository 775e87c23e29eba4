"""GPU: the one-call front-end (sdpl_frontend_process) equals the three separate extractors / the oracle, keeps the
previous frame across calls, and the N-rank sharding of a frame batch gives byte-identical per-frame results."""
import numpy as np
import pytest

from sdpl_slam_b200 import synth

pytestmark = pytest.mark.gpu


def _seq(n, h=240, w=416):
    out = []
    for s in range((n + 1) // 2):
        out += [synth.frame(500 + s, h, w), synth.partner(500 + s, h, w)]
    return np.stack(out[:n])


def test_frontend_matches_separate_calls_and_oracle(frontend, oracle):
    imgs = _seq(6)
    fe = frontend.FrontEnd(500, 1.2, 8, 20, 7, 0, 2, 0.8, 2, 2.0, 0.8, 64)
    r = fe.process(imgs)
    st = r["stats"]
    orb = frontend.ORBextractor(500, 1.2, 8, 20, 7)
    line = frontend.Lineextractor(0, 2, 0.8, 2, 2.0, 0)
    prev_d = prev_ld = None
    for f in range(len(imgs)):
        k, d = orb(imgs[f]); kl, dl = line(imgs[f])
        assert st["n_kp"][f] == len(k) and st["n_lines"][f] == len(kl)
        assert r["kps"][f, :len(k)].tobytes() == k.tobytes() and (r["desc"][f, :len(k)] == d).all()
        assert r["kls"][f, :len(kl)].tobytes() == kl.tobytes() and (r["ldesc"][f, :len(kl)] == dl).all()
        if f == 0:
            assert st["n_pt_matches"][0] == 0 and st["n_ln_matches"][0] == 0
            assert (r["pt_matches"][0, :len(k)]["train"] == -1).all()
        else:
            ref, nacc = oracle.match_ratio(d, prev_d, 0.8, 64)
            assert st["n_pt_matches"][f] == nacc
            np.testing.assert_array_equal(r["pt_matches"][f, :len(k)]["train"], ref["train"])
            np.testing.assert_array_equal(r["pt_matches"][f, :len(k)]["distance"], ref["distance"])
            if len(dl) and len(prev_ld):
                refl, naccl = oracle.match_ratio(dl, prev_ld, 0.8, 64)
                assert st["n_ln_matches"][f] == naccl
                np.testing.assert_array_equal(r["ln_matches"][f, :len(kl)]["train"], refl["train"])
        prev_d, prev_ld = d, dl
    # odd frames are shifted copies of the even ones: most keypoints find their partner
    assert st["n_pt_matches"][1] > 0.4 * st["n_kp"][1]


def test_frontend_keeps_previous_frame_across_calls(frontend):
    imgs = _seq(4)
    a = frontend.FrontEnd(500, 1.2, 8, 20, 7)
    whole = a.process(imgs)
    w_stats = whole["stats"].copy(); w_pm = whole["pt_matches"].copy()
    b = frontend.FrontEnd(500, 1.2, 8, 20, 7)
    s1 = b.process(imgs[:2])["stats"].copy()
    r2 = b.process(imgs[2:])
    np.testing.assert_array_equal(np.concatenate([s1, r2["stats"]]), w_stats)
    n = w_stats["n_kp"][2]
    np.testing.assert_array_equal(r2["pt_matches"][0, :n]["train"], w_pm[2, :n]["train"])
    b.reset()
    assert b.process(imgs[2:])["stats"]["n_pt_matches"][0] == 0


def test_sharded_batch_equals_single_rank(frontend):
    """BASELINE configs[3] in miniature: contiguous shards of a frame batch processed by independent handles give the same
    per-frame keypoints / keylines as the un-sharded batch (frames are independent units)."""
    imgs = _seq(8)
    orb = frontend.ORBextractor(500, 1.2, 8, 20, 7); line = frontend.Lineextractor()
    full_o = orb.extract_batch(imgs); full_l = line.extract_batch(imgs)
    for world in (2, 4):
        per = len(imgs) // world
        for rank in range(world):
            o2 = frontend.ORBextractor(500, 1.2, 8, 20, 7); l2 = frontend.Lineextractor()
            so = o2.extract_batch(imgs[rank * per:(rank + 1) * per]); sl = l2.extract_batch(imgs[rank * per:(rank + 1) * per])
            for j in range(per):
                f = rank * per + j
                assert so[j][0].tobytes() == full_o[f][0].tobytes() and so[j][1].tobytes() == full_o[f][1].tobytes()
                assert sl[j][0].tobytes() == full_l[f][0].tobytes() and sl[j][1].tobytes() == full_l[f][1].tobytes()


def test_handles_survive_geometry_changes(frontend):
    """One handle, alternating image sizes and batch sizes: results equal those of fresh handles."""
    orb = frontend.ORBextractor(500, 1.2, 8, 20, 7); line = frontend.Lineextractor()
    for (h, w, n) in ((240, 416, 1), (188, 320, 3), (240, 416, 2), (375, 1242, 1), (188, 320, 1)):
        imgs = np.stack([synth.frame(900 + i, h, w) for i in range(n)])
        a = orb.extract_batch(imgs); b = line.extract_batch(imgs)
        fo = frontend.ORBextractor(500, 1.2, 8, 20, 7); fl = frontend.Lineextractor()
        for i in range(n):
            k, d = fo(imgs[i]); kl, dl = fl(imgs[i])
            assert a[i][0].tobytes() == k.tobytes() and a[i][1].tobytes() == d.tobytes()
            assert b[i][0].tobytes() == kl.tobytes() and b[i][1].tobytes() == dl.tobytes()


def test_pipelined_submit_collect_equals_process(frontend):
    imgs = _seq(6)
    a = frontend.FrontEnd(500, 1.2, 8, 20, 7)
    ref = [dict((k, v.copy()) for k, v in a.process(imgs[i:i + 2]).items()) for i in (0, 2, 4)]
    b = frontend.FrontEnd(500, 1.2, 8, 20, 7)
    b.submit(imgs[0:2]); b.submit(imgs[2:4])
    out = [dict((k, v.copy()) for k, v in b.collect().items())]
    b.submit(imgs[4:6])
    out.append(dict((k, v.copy()) for k, v in b.collect().items()))
    out.append(dict((k, v.copy()) for k, v in b.collect().items()))
    for r, o in zip(ref, out):
        np.testing.assert_array_equal(r["stats"], o["stats"])
        for f in range(2):
            nk, nl = r["stats"]["n_kp"][f], r["stats"]["n_lines"][f]
            assert r["kps"][f, :nk].tobytes() == o["kps"][f, :nk].tobytes()
            assert r["desc"][f, :nk].tobytes() == o["desc"][f, :nk].tobytes()
            assert r["kls"][f, :nl].tobytes() == o["kls"][f, :nl].tobytes()
            assert r["pt_matches"][f, :nk].tobytes() == o["pt_matches"][f, :nk].tobytes()
            assert r["ln_matches"][f, :nl].tobytes() == o["ln_matches"][f, :nl].tobytes()
    # nothing in flight any more: collect must fail loudly
    b._inflight = [imgs[0:2]]
    with pytest.raises(frontend.SdplError):
        b.collect()


def test_line_capacity_overflow_is_reported_and_recoverable(frontend):
    """More key lines than kl_capacity: collect reports SDPL_ERR_CAPACITY (never a silent truncation), the matchers stay inside
    the capacity-sized blocks, the flag does not stick to later batches, and a bigger capacity gives the normal result."""
    imgs = _seq(4, 375, 1242)
    ref = frontend.FrontEnd(500, 1.2, 8, 20, 7).process(imgs)
    want = ref["stats"].copy()
    assert want["n_lines"].min() > 64
    fe = frontend.FrontEnd(500, 1.2, 8, 20, 7)
    fe.set_line_capacity(64)
    for _ in range(2):
        with pytest.raises(frontend.SdplError) as e:
            fe.process(imgs)
        assert e.value.code == frontend.SDPL_ERR_CAPACITY
    fe.set_line_capacity(2048)
    got = fe.process(imgs)
    np.testing.assert_array_equal(got["stats"]["n_lines"], want["n_lines"])
    np.testing.assert_array_equal(got["stats"]["n_kp"], want["n_kp"])
    np.testing.assert_array_equal(got["stats"]["n_pt_matches"][1:], want["n_pt_matches"][1:])
    np.testing.assert_array_equal(got["stats"]["n_ln_matches"][1:], want["n_ln_matches"][1:])
    with pytest.raises(frontend.SdplError):
        frontend.FrontEnd(500, 1.2, 8, 20, 7).collect()


def test_pipelined_batches_of_varying_size(frontend):
    """submit / collect with two batches in flight and batch sizes that change: results equal the one-call path, and a
    collected result stays valid until the collect after the next one."""
    imgs = _seq(10)
    one = frontend.FrontEnd(500, 1.2, 8, 20, 7).process(imgs)["stats"].copy()
    fe = frontend.FrontEnd(500, 1.2, 8, 20, 7)
    fe.submit(imgs[:2]); fe.submit(imgs[2:6])
    a = fe.collect(); a_stats = a["stats"].copy()
    fe.submit(imgs[6:])
    b = fe.collect()
    assert (a["stats"] == a_stats).all()          # still intact after the next collect
    c = fe.collect()
    np.testing.assert_array_equal(np.concatenate([a_stats, b["stats"], c["stats"]]), one)
