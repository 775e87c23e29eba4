"""One pass of every product kernel over F frames of 1242x375 (ORB, lines, both matchers, Frame post-processing) after one warm-up
pass -- the workload `ncu --set full` captures for profiles/ (second half of the launches = the warm pass).
Usage: python tools/prof_all.py F"""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from sdpl_slam_b200 import frontend as fe, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H, W = 375, 1242
dev = torch.device("cuda", 0)
imgs = torch.from_numpy(synth.sequence(0, F, H, W, workers=8)).to(dev)
planes = [synth.scene_planes(s, H, W) for s in range(4)]
rep = (F + 3) // 4
dm = torch.from_numpy(np.stack([p[0] for p in planes])).to(dev).repeat(rep, 1, 1)[:F].contiguous()
dd = torch.from_numpy(np.stack([p[1] for p in planes])).to(dev).repeat(rep, 1, 1)[:F].contiguous()
df = torch.from_numpy(np.stack([p[2] for p in planes])).to(dev).repeat(rep, 1, 1, 1)[:F].contiguous()
orb = fe.ORBextractor(2000, 1.2, 8, 20, 7); line = fe.Lineextractor(0, 2, 0.8, 2, 2.0, 0); m = fe.BinaryDescriptorMatcher(); post = fe.FramePost()
KC, LC = orb.max_keypoints(), 2048
u8, i32, f32 = torch.uint8, torch.int32, torch.float32
z = lambda *s, dt=u8: torch.zeros(s, dtype=dt, device=dev)
kps, desc, nk = z(F + 1, KC, 28), z(F + 1, KC, 32), z(F + 1, dt=i32)
kls, ldesc, nl = z(F + 1, LC, 68), z(F + 1, LC, 32), z(F + 1, dt=i32)
best, second, out, nacc = z(F, KC, 16), z(F, KC, 16), z(F, KC, 16), z(F, dt=i32)
cap = ((H + 3) // 4) * ((W + 3) // 4)
ok, oc, ofn, odp, olb, on = z(F, cap, 28), z(F, cap, 28), z(F, cap, 2, dt=f32), z(F, cap, dt=f32), z(F, cap, dt=i32), z(F, dt=i32)
pst, pcor, pfn, pdep, pidx, npt = z(F, KC, 28), z(F, KC, 28), z(F, KC, 2, dt=f32), z(F, KC, dt=f32), z(F, KC, dt=i32), z(F, dt=i32)
cs, items = z(F, 64 * 48 + 1, dt=i32), z(F, KC, dt=i32)
for _ in range(2):
    orb.extract_batch_dev(imgs.data_ptr(), F, W, H, kps.data_ptr() + KC * 28, desc.data_ptr() + KC * 32, KC, nk.data_ptr() + 4, sync=True)
    line.extract_batch_dev(imgs.data_ptr(), F, W, H, kls.data_ptr() + LC * 68, ldesc.data_ptr() + LC * 32, LC, nl.data_ptr() + 4, sync=True)
    m.knn2_batch_dev(desc.data_ptr() + KC * 32, nk.data_ptr() + 4, KC * 32, desc.data_ptr(), nk.data_ptr(), KC * 32, F, KC, KC, best.data_ptr(), second.data_ptr(), True)
    m.ratio_batch_dev(best.data_ptr(), second.data_ptr(), nk.data_ptr() + 4, F, KC, 0.8, 64, out.data_ptr(), nacc.data_ptr(), True)
    post.sample_objects_dev(dm.data_ptr(), dd.data_ptr(), df.data_ptr(), F, W, H, 4, 25.0, ok.data_ptr(), oc.data_ptr(), ofn.data_ptr(), odp.data_ptr(),
                            olb.data_ptr(), cap, on.data_ptr(), True)
    post.point_corres_dev(dm.data_ptr(), dd.data_ptr(), df.data_ptr(), F, W, H, kps.data_ptr() + KC * 28, nk.data_ptr() + 4, KC, 40.0, pst.data_ptr(),
                          pcor.data_ptr(), pfn.data_ptr(), pdep.data_ptr(), pidx.data_ptr(), npt.data_ptr(), True)
    post.grid_dev(F, W, H, kps.data_ptr() + KC * 28, nk.data_ptr() + 4, KC, cs.data_ptr(), items.data_ptr(), 64, 48, True)
print("ok", int(nk[1:].sum()), int(nl[1:].sum()), int(on.sum()))
