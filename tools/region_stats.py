"""Region-growing workload statistics from the oracle's trace (CPU only): for every seed the sequential LSD processes, its
position in the seed order, the pixels it expands and its first region size; then a model of K-seed waves (the GPU
schedule): which seeds share a wave, the longest region of each wave / of each 32-lane warp.  Guides the kernel design."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as orc
from sdpl_slam_b200 import synth

L = orc.lib()
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 3
img = synth.frame(seed, 375, 1242)
L.orc_lsd_trace(1)
lines = orc.lsd_detect(img)
buf = np.zeros(3 * 400000, np.int32)
n = L.orc_lsd_trace_get(buf.ctypes.data_as(C.c_void_p), len(buf))
tr = buf[:n].reshape(-1, 3)
L.orc_lsd_trace(0)
pos, exp, n1 = tr[:, 0], tr[:, 1], tr[:, 2]
print("seeds processed", len(tr), "lines", len(lines), "pixels expanded", exp.sum(), "of", 994 * 300)
for thr in (1, 4, 13, 32, 64, 128, 256, 512, 1024):
    m = exp >= thr
    print("  expanded >= %4d: %6d seeds, %8d expansions (%.1f %%)" % (thr, m.sum(), exp[m].sum(), 100 * exp[m].sum() / exp.sum()))
for K in (64, 128, 256):
    # waves: K consecutive processed seeds (a lower bound on seeds per wave: dead seeds of a wave are not in the trace)
    nw = (len(tr) + K - 1) // K
    crit = 0; crit_warp = 0; lane_sum = 0
    for w in range(nw):
        e = exp[w * K:(w + 1) * K]
        crit += e.max()
        lane_sum += e.sum()
        for j in range(0, len(e), 32):
            crit_warp += e[j:j + 32].max()
    print("K=%3d waves %4d  sum of wave maxima %8d  sum of warp maxima %8d (avg per warp %8d)  total %8d" %
          (K, nw, crit, crit_warp, crit_warp * 32 // K, lane_sum))
# how concentrated is the critical path: top regions
order = np.argsort(-exp)
print("largest regions:", exp[order[:20]].tolist())
print("their seed positions:", pos[order[:20]].tolist())

# ---- frontier model: steps needed when B list pixels are expanded per step (one region at a time) ----
L.orc_lsd_trace(2)
orc.lsd_detect(img)
buf = np.zeros(2000000, np.int32)
n = L.orc_lsd_trace_n(buf.ctypes.data_as(C.c_void_p), len(buf))
tn = buf[:n]
L.orc_lsd_trace(0)
growths = np.split(tn, np.where(tn == -1)[0])
growths = [g[g >= 0] for g in growths if (g >= 0).any()]
print("growths", len(growths), "expansions", sum(len(g) for g in growths))
for B in (1, 2, 4, 8, 16, 32):
    steps = 0
    for g in growths:
        i = 0; nn = 1
        while i < len(g):
            m = min(B, nn - i)
            i += m
            nn = g[i - 1]
            steps += 1
    print("B=%2d: %7d steps (%.2f expansions per step)" % (B, steps, sum(len(g) for g in growths) / steps))
for T in (8, 16, 32):
    big = [g for g in growths if len(g) > T]
    ex = sum(len(g) - T for g in big)
    for B in (4, 8, 16):
        steps = 0
        for g in big:
            i = T; nn = g[T - 1]
            while i < len(g):
                m = min(B, nn - i)
                i += m
                nn = g[i - 1]
                steps += 1
        print("growths longer than %2d: %5d, %6d expansions beyond; B=%2d: %6d steps (%.2f per step)" % (T, len(big), ex, B, steps, ex / max(1, steps)))
