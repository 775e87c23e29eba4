"""Where does a lines-only step of bench.py spend its time?  CUDA events after every part of the line stream."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
import bench
F = 512; H, W = bench.H, bench.W; CAP = bench.LINE_CAP
host = np.empty((F, H, W), np.uint8)
for p in range(F // 2):
    host[2 * p] = synth.frame(p, H, W); host[2 * p + 1] = synth.partner(p, H, W)
imgs = torch.from_numpy(host).cuda()
u8, i32 = torch.uint8, torch.int32
C = bench.LINE_CFG
line = fe.Lineextractor(C["nfeatures"], C["refine"], C["lsd_scale"], C["nlevels"], C["scale"], C["extractor"])
lmat = fe.BinaryDescriptorMatcher()
s = torch.cuda.Stream(priority=-1); s2 = torch.cuda.Stream()
line.set_stream(s.cuda_stream); lmat.set_stream(s2.cuda_stream)
d_kls = torch.empty((F, CAP, 68), dtype=u8, device='cuda'); d_ldesc = torch.zeros((F + 1, CAP, 32), dtype=u8, device='cuda')
d_nkl = torch.zeros(F + 1, dtype=i32, device='cuda')
best = torch.empty((F, CAP, 16), dtype=u8, device='cuda'); second = torch.empty((F, CAP, 16), dtype=u8, device='cuda')
nacc = torch.zeros(F, dtype=i32, device='cuda')
fs = CAP * 32
def step(ev):
    t0 = time.perf_counter()
    ev[0].record(s)
    line.extract_batch_dev(imgs.data_ptr(), F, W, H, d_kls.data_ptr(), d_ldesc.data_ptr() + fs, CAP, d_nkl.data_ptr() + 4)
    ev[1].record(s)
    t1 = time.perf_counter()
    s2.wait_stream(s)
    lmat.knn2_batch_dev(d_ldesc.data_ptr() + fs, d_nkl.data_ptr() + 4, fs, d_ldesc.data_ptr(), d_nkl.data_ptr(), fs, F, CAP, CAP, best.data_ptr(), second.data_ptr(), False)
    ev[2].record(s2)
    lmat.ratio_batch_dev(best.data_ptr(), second.data_ptr(), d_nkl.data_ptr() + 4, F, CAP, 0.8, 64, 0, nacc.data_ptr(), False)
    ev[3].record(s2)
    t2 = time.perf_counter()
    s.wait_stream(s2)
    return t1 - t0, t2 - t1
evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(6)]
for k in range(2): step(evs[0])
torch.cuda.synchronize()
t0 = time.perf_counter()
host_t = [step(evs[k]) for k in range(6)]
torch.cuda.synchronize()
print("wall per step ms", round((time.perf_counter() - t0) / 6 * 1e3, 1))
for k in range(1, 6):
    e = evs[k]
    print("step", k, "line", round(e[0].elapsed_time(e[1]), 1), "knn2", round(e[1].elapsed_time(e[2]), 1), "ratio", round(e[2].elapsed_time(e[3]), 2),
          "gap from prev end", round(evs[k - 1][3].elapsed_time(e[0]), 2), "host ms (line call, match calls)", [round(x * 1e3, 2) for x in host_t[k]])
