import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
seed, h, w, o = [int(x) for x in sys.argv[1:5]]
img = synth.frame(seed, h, w)
gs = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); gs.set_serial(True); gs(img); ps = gs.pending_rects(o)
gp = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); gp(img); pp = gp.pending_rects(o)
print("pending", len(ps), len(pp))
n = min(len(ps), len(pp))
d = np.abs(ps[:n, :5] - pp[:n, :5]).max(1)
bad = np.nonzero(d > 0)[0]
print("first differing", bad[:10])
for i in bad[:1]:
    for j in range(max(0, i - 2), min(n, i + 4)):
        t = int(pp[j, 7]); print(j, "S", np.round(ps[j, :5], 1), "seed", int(ps[j,5]) % 994, int(ps[j,5]) // 994, "n", int(ps[j,6]), "P", np.round(pp[j, :5], 1), "seed", int(pp[j,5]) % 994, int(pp[j,5]) // 994, "n", int(pp[j,6]), "wave", t >> 8, "lane", (t >> 1) & 127, "redo" if t & 1 else "good")
    print()
