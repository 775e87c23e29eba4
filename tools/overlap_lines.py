"""Two Lineextractor handles on two streams: does the tail of one batch's region growing overlap the next batch?"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H, W, CAP = 375, 1242, 2048
imgs = torch.from_numpy(synth.frames(range(F), H, W)).cuda()
u8, i32 = torch.uint8, torch.int32
hs = []
for i in range(2):
    g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0)
    s = torch.cuda.Stream(priority=-1)
    g.set_stream(s.cuda_stream)
    out = (torch.empty((F, CAP, 68), dtype=u8, device='cuda'), torch.empty((F, CAP, 32), dtype=u8, device='cuda'), torch.zeros(F, dtype=i32, device='cuda'))
    hs.append((g, s, out))
def run(k, which):
    g, s, o = hs[which]
    g.extract_batch_dev(imgs.data_ptr(), F, W, H, o[0].data_ptr(), o[1].data_ptr(), CAP, o[2].data_ptr())
for k in range(2): run(k, 0); run(k, 1)
torch.cuda.synchronize()
for name, sel in (("one handle", lambda k: 0), ("two handles", lambda k: k & 1)):
    t0 = time.perf_counter()
    for k in range(6): run(k, sel(k))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 6
    print(name, "ms per batch", round(dt * 1e3, 1), "lines fps", round(F / dt))
