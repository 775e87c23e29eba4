"""GPU: stage times of the EDLines back-end (extractor == 1) for batches of 1242x375 frames, beside the oracle on one host core.
Usage: python tools/prof_ed.py F [F ...]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
from oracle import oracle as orc
o = orc.LineOracle(0, 2, 0.8, 2, 2.0, 1)
img = synth.frame(0, 375, 1242)
t0 = time.perf_counter()
for _ in range(5): o(img)
print("oracle (one host core): %.1f ms per frame" % ((time.perf_counter() - t0) / 5 * 1e3), flush=True)
for F in map(int, sys.argv[1:]):
    imgs = synth.frames(range(F), 375, 1242)
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 1); g.set_profiling(True)
    if F == 1:
        run = lambda: g(imgs[0], capacity=4096)
    else:
        run = lambda: g.extract_batch(imgs, capacity=4096)
    run()
    t0 = time.perf_counter(); r = run(); dt = time.perf_counter() - t0
    st = {n: round(ms, 3) for n, ms, _ in g.stage_times()}
    nl = len(r[0]) if F == 1 else np.mean([len(k) for k, d in r])
    print("F=%d wall %.2f ms (%.3f ms per frame), mean lines %.1f, stages %s" % (F, dt * 1e3, dt * 1e3 / F, nl, st), flush=True)
