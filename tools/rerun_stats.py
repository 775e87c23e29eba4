import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
img = synth.sequence(0, 1, 375, 1242)[0]
g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True); g.grow_detail(2)
g(img, capacity=4096); g(img, capacity=4096)
p = g.grow_profile(0, 0)
print("cycles: taken-check %d, pipeline %d (growth %d), seeds already taken %d;" % (p["phaseA"], p["phaseB_busy"], p["phaseB_wait"], p["phaseA_wait"]))
print("reruns", p["reruns"], "first-growth pixels", p["grow"], "regrown+final pixels", p["rect_fit"], "rectangles", p["refine_tau"], "largest first region", p["grow_steps"],
      "rerun ms %.2f" % (p["rerun"] / 1.965e6))
