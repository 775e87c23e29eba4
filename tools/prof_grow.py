import sys, numpy as np, time
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
imgs = synth.frames(range(F), 375, 1242)
g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True); g.set_serial(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
g.extract_batch(imgs); t = time.time(); r = g.extract_batch(imgs); dt = time.time() - t
print("F", F, "host-api time", round(dt * 1e3, 1), "ms", "lines", [len(k) for k, d in r][:4])
print([(n, round(ms, 3)) for n, ms, _ in g.stage_times()])
for o in range(2):
    p = g.grow_profile(o, 0)
    tot = p["select"] + p["speculate"] + p["commit"] + p["rerun"]
    print("octave", o, {k: (round(v / 1.9e3, 1) if k in ("select", "speculate", "commit", "rerun") else v) for k, v in p.items()}, "total us", round(tot / 1.9e3))
