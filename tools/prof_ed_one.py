"""GPU: two passes of the EDLines back-end over one frame (for ncu -k regex:k_ed_ captures)."""
import sys
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 1)
img = synth.frame(0, 375, 1242)
for _ in range(2):
    k, d = g(img, capacity=4096)
print("ok", len(k))
