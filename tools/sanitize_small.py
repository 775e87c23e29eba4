import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
for (h, w) in ((96, 160), (211, 333)):
    img = synth.frame(5, h, w)
    for mode in (0, 1, 2, 3):
        g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_serial(mode)
        k, d = g(img)
        print("line", (h, w), "mode", mode, len(k))
    o = fe.ORBextractor(300, 1.2, 6, 20, 7)
    k, d = o(img); k2, d2 = o(synth.partner(5, h, w))
    m = fe.BinaryDescriptorMatcher(); b, s = m.knnMatch(d, d2); print("orb", len(k), "matches", int((b["train"] >= 0).sum()))
    c, r = m.radiusMatch(d, d2, 60, k=3)
f = fe.FrontEnd(300, 1.2, 6, 20, 7)
r = f.process(np.stack([synth.frame(6, 120, 200), synth.partner(6, 120, 200), synth.frame(7, 120, 200)]))
print(r["stats"])
