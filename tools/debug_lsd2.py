import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
gs = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); gs.set_serial(True)
gp = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); gp.set_serial(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
for seed in range(1, 13):
    h, w = (375, 1242) if seed % 3 else (480, 640)
    img = synth.frame(seed, h, w)
    ks, ds = gs(img)
    segs_s = [gs.lsd_segments(o) for o in range(2)]
    out = []
    for rep in range(2):
        kp, dp = gp(img)
        segs_p = [gp.lsd_segments(o) for o in range(2)]
        nd = []
        for o in range(2):
            if len(segs_p[o]) != len(segs_s[o]): nd.append(("len", len(segs_p[o]), len(segs_s[o])))
            else: nd.append(int((np.abs(segs_p[o] - segs_s[o]).max(1) > 0).sum()) if len(segs_s[o]) else 0)
        out.append(nd)
    print(seed, (h, w), [len(s) for s in segs_s], out)
