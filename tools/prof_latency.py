"""GPU tuning: single-frame latency of the line pipeline by stage, for (warps per task, CTAs per SM, phase-A cap) settings.
Usage: python tools/prof_latency.py "nw,mb,ta" ...   (no argument: the default schedule)"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
imgs = synth.sequence(0, 8, 375, 1242)
for spec in (sys.argv[1:] or ["default"]):
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True); g.grow_detail(True)
    if spec != "default":
        nw, mb, ta = map(int, spec.split(","))
        g.set_serial(0 | ((nw | (mb << 4)) << 8) | ((ta + 1) << 24))
    acc = {}
    ts = []
    for i in range(24):
        t0 = time.perf_counter(); g(imgs[i % 8], capacity=4096); dt = time.perf_counter() - t0
        if i >= 8:
            ts.append(dt)
            for n, ms, _ in g.stage_times():
                acc[n] = acc.get(n, 0) + ms / 16
    print(spec, "wall ms %.2f" % (1000 * np.median(ts)), {k: round(v, 2) for k, v in acc.items()}, g.grow_profile(0, 0), flush=True)
