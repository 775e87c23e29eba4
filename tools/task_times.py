"""GPU tuning: distribution of the region-growing task times over a batch (cycles of thread 0 from first to last wave) and how well
the number of seed candidates predicts them.  Usage: python tools/task_times.py F"""
import sys, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
imgs = synth.sequence(0, F, 375, 1242, workers=8)
g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True)
g.extract_batch(imgs, capacity=4096); g.extract_batch(imgs, capacity=4096)
st = dict((n, ms) for n, ms, _ in g.stage_times())
for o in (0, 1):
    t = np.array([sum(g.grow_profile(o, f)[k] for k in ("select", "speculate", "commit", "rerun")) for f in range(F)]) / 1.965e6
    seeds = np.array([g.grow_profile(o, f)["seeds"] for f in range(F)])
    waves = np.array([g.grow_profile(o, f)["waves"] for f in range(F)])
    print("octave", o, "task ms: min %.1f p10 %.1f median %.1f mean %.1f p90 %.1f p99 %.1f max %.1f" %
          (t.min(), np.percentile(t, 10), np.median(t), t.mean(), np.percentile(t, 90), np.percentile(t, 99), t.max()),
          "| corr(time, seeds) %.3f corr(time, waves) %.3f" % (np.corrcoef(t, seeds)[0, 1], np.corrcoef(t, waves)[0, 1]))
    order = np.argsort(-seeds)
    print("   the 5 %% of tasks with most seeds hold %d of the 26 slowest tasks" % len(set(order[:F // 20]) & set(np.argsort(-t)[:26])))
print("grow kernel ms", st["lsd_grow"])
