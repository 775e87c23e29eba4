"""GPU diagnostic: which region-growing schedules disagree with the oracle on a batch of frames (line counts per frame)."""
import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
from oracle import oracle as orc
F = int(sys.argv[1]) if len(sys.argv) > 1 else 96
seeds = range(1000, 1000 + F)
imgs = synth.frames(seeds, 375, 1242)
ref = orc.LineOracle(0, 2, 0.8, 2, 2.0, 0)
want = [len(ref(im)[0]) for im in imgs]
def run(mode, batch):
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0)
    if mode is not None:
        g.set_serial(mode)
    out = []
    for b in range(0, F, batch):
        out += [len(k) for k, d in g.extract_batch(imgs[b:b + batch], capacity=4096)]
    return out
for name, mode, batch in (("auto96", None, F), ("auto96-again", None, F), ("auto8", None, 8), ("seq(1) batch8", 1, 8), ("4x4 batch96", (4 | (4 << 4)) << 8, F),
                          ("4x4 batch8", (4 | (4 << 4)) << 8, 8), ("8x1 batch96", (8 | (1 << 4)) << 8, F), ("2x8 batch96", (2 | (8 << 4)) << 8, F)):
    got = run(mode, batch)
    bad = [(i, got[i], want[i]) for i in range(F) if got[i] != want[i]]
    print(name, "mismatching frames:", bad)
