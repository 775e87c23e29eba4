"""GPU diagnostic: rectangles handed to the NFA stage and their verdicts, GPU vs oracle, octave 0 of a frame."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
from oracle import oracle as orc
L = orc.lib()
for seed in [int(x) for x in sys.argv[1:]]:
    img = synth.frame(seed, 375, 1242)
    L.orc_lsd_trace(1)
    seg = orc.lsd_detect(img)
    buf = np.zeros(9 * 20000); n = L.orc_lsd_trace_cand(buf.ctypes.data_as(C.c_void_p), len(buf))
    L.orc_lsd_trace(0)
    oc = buf[:n].reshape(-1, 9)
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0)
    k, d = g(img)
    gp = g.pending_rects(0)
    print("seed", seed, "oracle rects", len(oc), "accepted", int((oc[:, 6] > 0).sum()), "| gpu rects", len(gp), "accepted", int(gp[:, 6].sum()), "| segments", len(seg), len(g.lsd_segments(0)))
    m = min(len(oc), len(gp))
    dif = np.abs(oc[:m, :5] - gp[:m, :5]).max(1)
    print("  max |rect diff| over common prefix", dif.max(), "first index with diff > 1e-6:", (np.where(dif > 1e-6)[0][:3]).tolist())
    acc_o = oc[:m, 6] > 0; acc_g = gp[:m, 6] > 0.5
    for i in np.where(acc_o != acc_g)[0][:5]:
        print("  rect", i, "oracle first/final log_nfa %.17g %.17g" % (oc[i, 5], oc[i, 6]), "n,k", oc[i, 7], oc[i, 8], "gpu accepted", gp[i, 6], "rect", oc[i, :5].tolist(), "gpu", gp[i, :5].tolist(), "tag", gp[i, 7])
