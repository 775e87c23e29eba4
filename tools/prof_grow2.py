import sys, numpy as np, time
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1])
imgs = synth.frames(range(F), 375, 1242)
for mode in [int(x) for x in sys.argv[2:]]:
  g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True); g.set_serial(mode)
  g.extract_batch(imgs); r = g.extract_batch(imgs)
  st = dict((n, ms) for n, ms, _ in g.stage_times())
  print({k: round(v, 2) for k, v in st.items()})
  for oc in (0, 1):
    p = g.grow_profile(0, oc)
    print("oct", oc, "F", F, "mode", mode & 3, "W", mode >> 8, "lines", sum(len(k) for k, d in r), "grow ms", round(st["lsd_grow"], 1), {k: (round(v / 1.9e3) if k in ("select", "speculate", "commit", "rerun") else v) for k, v in p.items()})
