"""GPU: two passes of the line front-end over F frames (first = warm-up), for `ncu -k regex:k_lsd_grow2 --launch-skip ...` source-level
captures of the region-growing kernel.  Usage: python tools/prof_one.py F"""
import sys
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1
imgs = synth.sequence(0, F, 375, 1242) if F > 1 else [synth.frame(0, 375, 1242)]
g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0)
for _ in range(2):
    if F == 1:
        k, d = g(imgs[0], capacity=4096)
    else:
        res = g.extract_batch(imgs, capacity=4096)
print("ok", F)
