"""GPU tuning: single-frame line-extraction latency for the SDPL_GROW_SMALL settings given on the command line ("warps,cap"), with a
check that the output does not change.  Usage: python tools/prof_latency2.py "8,16" "16,16" ..."""
import os, subprocess, sys
CODE = r'''
import sys, time, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
imgs = synth.sequence(0, 8, 375, 1242)
g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True)
ts = []; acc = {}; sig = []
for i in range(24):
    t0 = time.perf_counter(); k, d = g(imgs[i % 8], capacity=4096); dt = time.perf_counter() - t0
    if i < 8: sig.append(hash((k.tobytes(), d.tobytes())))
    if i >= 8:
        ts.append(dt)
        for n, ms, _ in g.stage_times(): acc[n] = acc.get(n, 0) + ms / 16
p = g.grow_profile(0, 0)
print("wall ms %.2f grow %.2f nfa %.2f" % (1000 * np.median(ts), acc["lsd_grow"], acc["lsd_nfa"]), {k: round(p[k] / 1.965e6, 2) for k in ("select", "speculate", "commit", "rerun")}, p["waves"], p["reruns"], "sig", hash(tuple(sig)) % 100000)
'''
for spec in sys.argv[1:]:
    env = dict(os.environ, SDPL_GROW_SMALL=spec)
    r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(spec, r.stdout.strip(), r.stderr.strip()[-300:], flush=True)
