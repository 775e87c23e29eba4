import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
from oracle import oracle as orc
seed, h, w = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1, 375, 1242)
img = synth.frame(seed, h, w)
ref = orc.LineOracle(0, 2, 0.8, 2, 2.0, 0); kr, dr = ref(img)
for serial in (1, 0):
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_serial(bool(serial))
    kg, dg = g(img)
    print("serial" if serial else "spec", "lines", len(kg), "ref", len(kr), "stages", [(n, round(t, 3)) for n, t, _ in g.stage_times()] if False else "")
    for o in range(2):
        sg, sr = g.lsd_segments(o), ref.last_segments(o)
        print(" octave", o, len(sg), len(sr))
        n = min(len(sg), len(sr))
        d = np.abs(sg[:n] - sr[:n]).max(1) if n else np.zeros(0)
        bad = np.nonzero(d > 0)[0]
        print("  differing:", len(bad), "max", d.max() if n else 0)
        for i in bad[:6]:
            print("   ", i, sg[i], sr[i])
