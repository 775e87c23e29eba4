"""GPU tuning: stage times of the line pipeline and the matcher for the occupancy knobs in the environment (SDPL_NFA1_MINB,
SDPL_MATCH_MINB, SDPL_LBD_MINB).  Usage: python tools/prof_occ.py F"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 512
imgs = synth.sequence(0, F, 375, 1242, workers=8)
g = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); g.set_profiling(True)
g.extract_batch(imgs, capacity=4096); g.extract_batch(imgs, capacity=4096)
st = dict((n, ms) for n, ms, _ in g.stage_times())
orb = fe.ORBextractor(2000, 1.2, 8, 20, 7); m = fe.BinaryDescriptorMatcher(); m.set_profiling(True)
d_imgs = torch.from_numpy(imgs).cuda()
KC = orb.max_keypoints()
kps = torch.zeros((F + 1, KC, 28), dtype=torch.uint8, device="cuda"); desc = torch.zeros((F + 1, KC, 32), dtype=torch.uint8, device="cuda")
nk = torch.zeros(F + 1, dtype=torch.int32, device="cuda"); best = torch.zeros((F, KC, 16), dtype=torch.uint8, device="cuda"); second = torch.zeros_like(best)
orb.extract_batch_dev(d_imgs.data_ptr(), F, 1242, 375, kps.data_ptr() + KC * 28, desc.data_ptr() + KC * 32, KC, nk.data_ptr() + 4, sync=True)
for _ in range(3):
    m.knn2_batch_dev(desc.data_ptr() + KC * 32, nk.data_ptr() + 4, KC * 32, desc.data_ptr(), nk.data_ptr(), KC * 32, F, KC, KC, best.data_ptr(), second.data_ptr(), True)
mt = dict((n, ms) for n, ms, _ in m.stage_times())
print({k: round(v, 3) for k, v in st.items() if k in ("lsd_nfa", "lbd_bands")}, {k: round(v, 3) for k, v in mt.items()}, "checksum", int(best.sum()) % 100003)
