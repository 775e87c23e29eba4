"""Deterministic synthetic frames shared by the oracle, the CPU baseline and the GPU path
(SURVEY.md section 8d).  Pure numpy integer arithmetic, seeded with PCG64, so the same seed gives the
same bytes on every machine.

frame(seed) = low-resolution uniform noise (1/4 resolution) bilinearly upsampled with integer weights
(gives FAST corners everywhere) + 40 filled random rectangles (gives LSD line segments) + +-2 grey
levels of pixel noise.  partner(seed) = the same frame shifted by a few integer pixels with a
brightness offset (for frame-to-frame matching).
"""
import numpy as np


def frame(seed: int, h: int = 375, w: int = 1242, n_rect: int = 40) -> np.ndarray:
    rng = np.random.default_rng(int(seed))
    gh, gw = (h + 3) // 4 + 2, (w + 3) // 4 + 2
    lat = rng.integers(0, 256, size=(gh, gw), dtype=np.int64)
    ys, xs = np.arange(h), np.arange(w)
    y0, fy = ys // 4, ys % 4
    x0, fx = xs // 4, xs % 4
    a = lat[y0][:, x0] * (4 - fx)[None, :] + lat[y0][:, x0 + 1] * fx[None, :]
    b = lat[y0 + 1][:, x0] * (4 - fx)[None, :] + lat[y0 + 1][:, x0 + 1] * fx[None, :]
    img = (a * (4 - fy)[:, None] + b * fy[:, None] + 8) // 16
    # compress the texture contrast a little and paint rectangles
    img = 64 + img // 2
    for _ in range(n_rect):
        rw = int(rng.integers(20, max(21, w // 4)))
        rh = int(rng.integers(20, max(21, h // 2)))
        rx = int(rng.integers(0, max(1, w - rw)))
        ry = int(rng.integers(0, max(1, h - rh)))
        g = int(rng.integers(0, 256))
        tex = int(rng.integers(0, 3))   # 0: flat, 1/2: keep some texture inside
        if tex == 0:
            img[ry:ry + rh, rx:rx + rw] = g
        else:
            img[ry:ry + rh, rx:rx + rw] = (img[ry:ry + rh, rx:rx + rw] + 3 * g) // 4
    img = img + rng.integers(-2, 3, size=(h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def partner_from(f: np.ndarray, dx: int = 3, dy: int = 1, gain: int = 6) -> np.ndarray:
    """Integer-shifted, brightness-offset copy of the frame `f` (edge pixels replicated)."""
    h, w = f.shape
    ys = np.clip(np.arange(h) - dy, 0, h - 1)
    xs = np.clip(np.arange(w) - dx, 0, w - 1)
    return np.clip(f.astype(np.int64)[ys][:, xs] + gain, 0, 255).astype(np.uint8)


def partner(seed: int, h: int = 375, w: int = 1242, dx: int = 3, dy: int = 1, gain: int = 6) -> np.ndarray:
    """partner_from(frame(seed))."""
    return partner_from(frame(seed, h, w), dx, dy, gain)


def _pair(args):
    seed, h, w = args
    f = frame(seed, h, w)
    return f, partner_from(f)


def sequence(g0: int, g1: int, h: int = 375, w: int = 1242, workers: int = 1) -> np.ndarray:
    """Frames [g0, g1) of THE synthetic sequence frame(0), partner(0), frame(1), partner(1), ...: global frame g is
    frame(g // 2) for even g and partner(g // 2) for odd g.  This is the batch BASELINE.json configs[3] shards across GPUs;
    every rank synthesises its own block (and the halo frame before it) from the global frame numbers alone."""
    out = np.empty((max(0, g1 - g0), h, w), np.uint8)
    if g1 <= g0:
        return out
    seeds = list(range(g0 // 2, (g1 + 1) // 2))
    if workers > 1 and len(seeds) > 4:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers) as pool:
            pairs = pool.map(_pair, [(s, h, w) for s in seeds], chunksize=4)
    else:
        pairs = [_pair((s, h, w)) for s in seeds]
    for s, (a, b) in zip(seeds, pairs):
        for g, img in ((2 * s, a), (2 * s + 1, b)):
            if g0 <= g < g1:
                out[g - g0] = img
    return out


def frames(seeds, h: int = 375, w: int = 1242) -> np.ndarray:
    return np.stack([frame(s, h, w) for s in seeds])


def map_descriptors(desc: np.ndarray, n_map: int, seed: int, flips: int = 12) -> np.ndarray:
    """Config-3 style map: the frame's own descriptors with `flips` random bits flipped, padded with
    uniform random 256-bit strings up to n_map rows, then shuffled."""
    rng = np.random.default_rng(int(seed))
    own = desc.copy()
    for i in range(own.shape[0]):
        bits = rng.choice(256, size=flips, replace=False)
        for b in bits:
            own[i, b >> 3] ^= np.uint8(1 << (b & 7))
    rest = rng.integers(0, 256, size=(max(0, n_map - own.shape[0]), 32), dtype=np.uint8)
    allm = np.concatenate([own, rest])[:n_map]
    return allm[rng.permutation(allm.shape[0])].copy()


def scene_planes(seed: int, h: int = 375, w: int = 1242, n_obj: int = 6):
    """The per-frame planes Frame::Frame reads next to the image (src/Frame.cc): a semantic / motion mask (int32, 0 =
    background, 1.. = object label: a few rectangles), a depth plane (float32 metres, with holes of 0 and a far region beyond
    the thresholds) and an optical-flow plane (float32, 2 channels; exactly 0 in some patches, pointing out of the image near
    two borders).  Deterministic in the seed."""
    rng = np.random.default_rng(int(seed) + 777_000)
    mask = np.zeros((h, w), np.int32)
    for k in range(n_obj):
        rw = int(rng.integers(w // 16, w // 5)); rh = int(rng.integers(h // 8, h // 2))
        rx = int(rng.integers(0, w - rw)); ry = int(rng.integers(0, h - rh))
        mask[ry:ry + rh, rx:rx + rw] = k + 1
    yy, xx = np.mgrid[0:h, 0:w]
    depth = (4.0 + 36.0 * (1.0 - yy / float(h)) + 2.0 * np.sin(xx / 37.0)).astype(np.float32)        # 4 .. 42 m, far at the top
    depth[mask > 0] = (6.0 + 3.0 * mask[mask > 0]).astype(np.float32)                                # objects: flat depth per label
    holes = rng.integers(0, 50, size=(h // 5 + 1, w // 5 + 1)) == 0
    depth[np.kron(holes, np.ones((5, 5), bool))[:h, :w]] = 0.0
    flow = np.empty((h, w, 2), np.float32)
    flow[..., 0] = (1.5 + xx / 300.0 - 2.0 * (mask > 0)).astype(np.float32)
    flow[..., 1] = (-0.75 + yy / 250.0 + 0.5 * (mask % 2)).astype(np.float32)
    zero = rng.integers(0, 30, size=(h // 8 + 1, w // 8 + 1)) == 0
    flow[np.kron(zero, np.ones((8, 8), bool))[:h, :w]] = 0.0
    flow[:, w - 3:, 0] += 6.0          # leaves the image on the right
    flow[:2, :, 1] -= 4.0              # and at the top
    return mask, depth, flow
