"""GPU tuning: NFA stage time for the register caps of the second pass.  Usage: python tools/prof_nfa.py F minb minb ..."""
import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1])
imgs = synth.sequence(0, F, 375, 1242, workers=8)
ref = None
for mb in map(int, sys.argv[2:]):
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True)
    g.set_serial(mb << 3)
    g.extract_batch(imgs, capacity=4096); r = g.extract_batch(imgs, capacity=4096)
    st = dict((n, ms) for n, ms, _ in g.stage_times())
    sig = [(k.tobytes(), d.tobytes()) for k, d in r]
    ref = ref or sig
    print("minb", mb, {k: round(v, 2) for k, v in st.items()}, "same:", sig == ref, flush=True)
