import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
imgs = synth.frames(range(F), 375, 1242)
o = fe.ORBextractor(2000, 1.2, 8, 20, 7); o.set_profiling(True)
o.extract_batch(imgs); r = o.extract_batch(imgs)
print("F", F, [(n, round(ms, 3)) for n, ms, _ in o.stage_times()], "kps", [len(k) for k, d in r][:3])
