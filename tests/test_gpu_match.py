"""GPU parity: brute-force Hamming matcher vs the oracle (bit-exact indices and distances)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_desc(rng, n):
    return rng.integers(0, 256, size=(n, 32), dtype=np.uint8)


def _noisy_copy(rng, d, flips):
    out = d.copy()
    for i in range(out.shape[0]):
        for b in rng.choice(256, size=flips, replace=False):
            out[i, b >> 3] ^= np.uint8(1 << (b & 7))
    return out


@pytest.mark.parametrize("nq,nt", [(2000, 2000), (1000, 5000), (1, 1), (3, 2), (33, 65), (257, 31), (8000, 8000)])
def test_knn2_parity(frontend, oracle, nq, nt):
    rng = np.random.default_rng(nq * 7 + nt)
    t = _rand_desc(rng, nt)
    q = _rand_desc(rng, nq)
    m = min(nq, nt) // 2
    if m:
        q[:m] = _noisy_copy(rng, t[rng.permutation(nt)[:m]], 10)
    g = frontend.BinaryDescriptorMatcher()
    bg, sg = g.knnMatch(q, t, 2)
    br, sr = oracle.match_knn2(q, t)
    for name in ("query", "train", "img", "distance"):
        np.testing.assert_array_equal(bg[name], br[name], err_msg="best." + name)
        np.testing.assert_array_equal(sg[name], sr[name], err_msg="second." + name)


def test_knn2_ties_go_to_lowest_index(frontend, oracle):
    rng = np.random.default_rng(2)
    base = _rand_desc(rng, 50)
    t = np.concatenate([base, base, base])          # every distance appears three times
    q = _noisy_copy(rng, base, 5)
    g = frontend.BinaryDescriptorMatcher()
    bg, sg = g.knnMatch(q, t, 2)
    br, sr = oracle.match_knn2(q, t)
    np.testing.assert_array_equal(bg["train"], br["train"])
    np.testing.assert_array_equal(sg["train"], sr["train"])
    assert (bg["train"] == np.arange(50)).all() and (sg["train"] == np.arange(50) + 50).all()


def test_knn2_single_train_has_no_second(frontend, oracle):
    rng = np.random.default_rng(3)
    q, t = _rand_desc(rng, 10), _rand_desc(rng, 1)
    bg, sg = frontend.BinaryDescriptorMatcher().knnMatch(q, t, 2)
    br, sr = oracle.match_knn2(q, t)
    np.testing.assert_array_equal(sg["train"], sr["train"])
    assert (sg["train"] == -1).all() and (sg["distance"] == 257).all()
    np.testing.assert_array_equal(bg["distance"], br["distance"])


def test_empty_inputs(frontend):
    g = frontend.BinaryDescriptorMatcher()
    b, s = g.knnMatch(np.zeros((0, 32), np.uint8), np.zeros((5, 32), np.uint8))
    assert len(b) == 0
    b, s = g.knnMatch(np.zeros((4, 32), np.uint8), np.zeros((0, 32), np.uint8))
    assert (b["train"] == -1).all()


def test_ratio_and_radius(frontend, oracle):
    rng = np.random.default_rng(4)
    t = _rand_desc(rng, 700)
    q = np.concatenate([_noisy_copy(rng, t[:300], 20), _rand_desc(rng, 100)])
    g = frontend.BinaryDescriptorMatcher()
    og, ng = g.ratioMatch(q, t, 0.8, 60)
    orr, nr = oracle.match_ratio(q, t, 0.8, 60)
    assert ng == nr and ng >= 290
    for name in ("query", "train", "distance"):
        np.testing.assert_array_equal(og[name], orr[name])
    cg, rg = g.radiusMatch(q, t, 100, k=5)
    cr, rr = oracle.match_radius(q, t, 100, 5)
    np.testing.assert_array_equal(cg, cr)
    np.testing.assert_array_equal(rg["train"], rr["train"])
    np.testing.assert_array_equal(rg["distance"], rr["distance"])


def test_hamming_properties_full_size(frontend):
    # size-independent properties at BASELINE size (8000 x 8000): d(x,x)=0, symmetry of the best distance
    rng = np.random.default_rng(6)
    d = _rand_desc(rng, 8000)
    g = frontend.BinaryDescriptorMatcher()
    b, s = g.knnMatch(d, d, 2)
    assert (b["train"] == np.arange(8000)).all() and (b["distance"] == 0).all()
    # second best of i is j  =>  distance(j -> i) is at most that distance
    j = s["train"]
    assert (s["distance"][j] <= s["distance"]).all()


def test_knn_general_k_and_stored_train_set(frontend, oracle):
    rng = np.random.default_rng(8)
    t1, t2 = _rand_desc(rng, 300), _rand_desc(rng, 200)
    q = _noisy_copy(rng, np.concatenate([t1[:40], t2[:40]]), 12)
    g = frontend.BinaryDescriptorMatcher()
    # k = 4 nearest == brute force over the full distance matrix (ties -> lowest index)
    allt = np.concatenate([t1, t2])
    r = g.knnMatch(q, allt, k=4)
    full = np.unpackbits(q[:, None, :] ^ allt[None, :, :], axis=2).sum(2)
    order = np.lexsort((np.broadcast_to(np.arange(full.shape[1]), full.shape), full), axis=1)[:, :4]
    np.testing.assert_array_equal(r["train"], order)
    np.testing.assert_array_equal(r["distance"], np.take_along_axis(full, order, 1))
    # add / train / match / knnMatch / radiusMatch on the stored (device-resident) set, through sdpl_matcher_add / _knn / _radius:
    # trainIdx = row of the concatenation, imgIdx = image, as the reference composes them (oracle.match_stored_knn)
    g.add([t1, t2]); g.train()
    assert g.train_size() == (500, 2)
    m = g.match(q)
    want1 = oracle.match_stored_knn(q, [t1, t2], 1)[:, 0]
    for name in ("query", "train", "img", "distance"):
        np.testing.assert_array_equal(m[name], want1[name])
    assert (m["img"][:40] == 0).all() and (m["img"][40:] == 1).all()
    assert (m["train"][:40] == np.arange(40)).all() and (m["train"][40:] == 300 + np.arange(40)).all()
    for k in (2, 5):
        r = g.knnMatch(q, k=k)
        r = np.stack(r, 1) if k == 2 else r
        want = oracle.match_stored_knn(q, [t1, t2], k)
        for name in ("train", "img", "distance"):
            np.testing.assert_array_equal(r[name], want[name])
    counts, near = g.radiusMatch(q, None, 100, k=3)
    wc, wn = oracle.match_radius(q, allt, 100, 3)
    np.testing.assert_array_equal(counts, wc); np.testing.assert_array_equal(near["train"], wn["train"])
    # the set grows across add() calls without losing what is stored
    g.add(t1[:7])
    assert g.train_size() == (507, 3)
    assert (g.match(t1[:7])["img"] == 0).all()            # ties (distance 0 twice) go to the lowest row
    g.clear()
    assert g.train_size() == (0, 0)
    with pytest.raises(ValueError):
        g.match(q)
