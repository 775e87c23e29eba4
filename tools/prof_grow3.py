"""GPU tuning: region-growing stage time of the two-phase schedule for (warps per task, CTAs per SM, phase-A cap) settings,
against the round-1 schedule, on F frames of 1242x375.  Usage: python tools/prof_grow3.py F "nw,mb,ta" "nw,mb,ta" ... ("legacy" = round 1)"""
import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1])
imgs = synth.frames(range(F), 375, 1242)
ref = None
for spec in sys.argv[2:]:
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True); g.grow_detail(True)
    if spec == "legacy":
        g.set_serial(4)
    else:
        nw, mb, ta = map(int, spec.split(","))
        g.set_serial(0 | ((nw | (mb << 4)) << 8) | ((ta + 1) << 24))
    g.extract_batch(imgs, capacity=4096); r = g.extract_batch(imgs, capacity=4096)
    st = dict((n, ms) for n, ms, _ in g.stage_times())
    sig = [(k.tobytes(), d.tobytes()) for k, d in r]
    if ref is None:
        ref = sig
    p = g.grow_profile(0, 0)
    print(spec, "grow ms %.2f" % st["lsd_grow"], "nfa %.2f" % st.get("lsd_nfa", 0), "same output as first:", sig == ref,
          {k: (v if k in ("waves", "reruns", "dead", "seeds", "grow_steps") else round(v / 1.9e3)) for k, v in p.items()}, flush=True)
