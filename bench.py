#!/usr/bin/env python
"""bench.py -- front-end frames/sec (ORB + LSD/LBD + Hamming match) at 1242x375 (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)
    python bench.py --config {1,2,3,5}                        # the other BASELINE.json configs (one JSON line each, profiles/)
    python bench.py --scaling strong --total-frames 4096      # configs[3] read literally: 4096 frames in total at every N
    python bench.py --sweep                                   # batch-size sweep 1..2048 through the host-buffer API

One "step" = one pass of the front-end over the rank's shard of THE synthetic sequence frame(0), partner(0), frame(1), ...
(sdpl_slam_b200/synth.py: sequence): rank r owns the contiguous block shard.shard_range gives it and also extracts the frame
before its block (shard.shard_with_halo), so that every frame of the batch -- the first frame of a shard included -- is matched
against its predecessor (ORB and LBD descriptors, knn-2 + ratio test).  `value` is timed with the frames already resident in
HBM; `e2e` goes through the host-buffer C ABI (H2D and D2H copies inside the timed region); `latency` is ONE frame per call
through sdpl_orb_extract + sdpl_line_extract + sdpl_match_ratio with pageable host buffers (what Frame::ExtractORB /
ExtractLines do, src/Frame.cc:927-949).  Prints ONE JSON line on rank 0.  The oracle under oracle/ is used here only for the
cpu_baseline / --impl reference / --verify legs (it is the CPU restatement of the reference, pinned to the reference's own
sources by tests/test_oracle_vs_ref.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LINE_CFG = dict(nfeatures=0, refine=2, lsd_scale=0.8, nlevels=2, scale=2.0, extractor=0)
RATIO, MAX_DIST = 0.8, 64
LINE_CAP = 2048
# BASELINE.json configs (0-based index in `baseline`); --config takes the survey's 1-based numbering
CONFIGS = {
    1: dict(baseline=0, w=1242, h=375, orb=dict(nfeatures=2000, scale=1.2, nlevels=8, ini=20, mn=7), line=None, match=None, frames=512,
            workload="ORBextractor on synthetic 1242x375 grayscale frames, 2000 features, 8 levels, scale 1.2 [BASELINE.json configs[0]]"),
    2: dict(baseline=1, w=1242, h=375, orb=dict(nfeatures=2000, scale=1.2, nlevels=8, ini=20, mn=7), line=LINE_CFG, match="prev", frames=512,
            workload="KITTI-size 1242x375 point+line front-end: ORB 2000 (8 levels, x1.2, FAST 20/7) + LSD/LBD lines (refine ADV, 0.8, 2 octaves) "
                     "+ frame-to-frame (t vs t-1) Hamming knn-2 ratio matching of ORB and LBD descriptors [BASELINE.json configs[1]]"),
    3: dict(baseline=2, w=640, h=480, orb=dict(nfeatures=1000, scale=1.2, nlevels=8, ini=20, mn=7), line=LINE_CFG, match="map", n_map=5000,
            frames=512,
            workload="TUM-size 640x480 frame stream: ORB 1000 + LSD/LBD lines, every frame's ORB descriptors matched (knn-2 + ratio) against "
                     "5000 map-point descriptors [BASELINE.json configs[2]]"),
    5: dict(baseline=4, w=2048, h=1536, orb=dict(nfeatures=8000, scale=1.2, nlevels=12, ini=20, mn=7), line=None, match="prev", frames=32,
            workload="2048x1536 high-res stress: ORB 8000 features, 12 levels, frame-to-frame 8000x8000 brute-force Hamming knn-2 + ratio test "
                     "[BASELINE.json configs[4]]"),
}
METRIC = "front-end frames/sec (ORB+LSD/LBD+match) at 1242x375"
UNIT = "frames/s"
# popc roofline of the matcher: popc.b32 issues at 16 / clk / SM; the kernel counts 4 words per 256-bit pair
# after the carry-save tree (8 without it)
POPC_PER_S = 16 * 148 * 1.965e9
POPC_PER_PAIR = 4


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _ncu_facts():
    """Per-stage facts that only a profiler can give (DRAM bytes per frame, issue-slot utilisation), written by
    tools/summarize_profiles.py from the `ncu --set full` captures of this build into profiles/ncu_stage_facts.json together
    with the capture they come from.  Absent file / stage -> null in the JSON line (never a stale constant in this file)."""
    p = os.path.join(ROOT, "profiles", "ncu_stage_facts.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def _level_sizes(w, h, scale, nlevels):
    sf = np.float32(1.0); out = []
    for l in range(nlevels):
        if l:
            sf = np.float32(sf * np.float32(scale))
        inv = np.float32(1.0) / sf
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def algorithmic_bytes(cfg):
    """Minimal per-frame traffic of each stage: compulsory reads + writes only."""
    w, h, oc = cfg["w"], cfg["h"], cfg["orb"]
    lv = _level_sizes(w, h, oc["scale"], oc["nlevels"])
    px = [a * b for a, b in lv]
    padded = [(a + 38) * (b + 38) for a, b in lv]
    nk = oc["nfeatures"]
    ncand = 7.5 * nk
    b = {}
    b["pyramid"] = w * h + sum(px[:-1]) + sum(padded)
    b["fast_score"] = sum(padded) + sum(px)
    b["cell_nms"] = 2 * sum(px) + 5 * ncand
    b["quadtree"] = ncand * 7 + nk * 5
    b["blur7"] = 2 * sum(px)
    b["orient_describe"] = nk * (749 + 512 + 5) + nk * (28 + 32)
    nt = cfg.get("n_map", nk)
    b["match_partial"] = (nk + nt) * 32 + nk * 16
    b["match_merge"] = nk * 16 + nk * 32
    # lines: level 0 = w x h, level 1 = round(w/2) x round(h/2); LSD works on the 0.8x blurred level
    l1 = (int(np.rint(w / 2)), int(np.rint(h / 2)))
    lsd_px = sum(int(np.rint(a * 0.8)) * int(np.rint(c * 0.8)) for a, c in [(w, h), l1])
    src_px = w * h + l1[0] * l1[1]
    b["lsd_pyramid"] = w * h + l1[0] * l1[1]
    b["lsd_scale"] = src_px + lsd_px
    b["lsd_gradient"] = lsd_px + lsd_px * (8 + 8 + 4)            # angle f64, cos/sin f32x2, grad u32
    b["lsd_sort"] = lsd_px * (2 + 4) * 2
    b["lsd_grow"] = lsd_px * (8 + 8 + 4 + 1 + 4)                  # every pixel's angle/trig/grad/used touched once + region list
    b["lsd_nfa"] = lsd_px * 8
    lbd_px = w * h + (w // 2) * (h // 2)
    b["lbd_sobel"] = w * h * 2 + lbd_px * (1 + 4)
    b["lbd_bands"] = 600 * (63 * 40 * 4 + 32)
    b["keylines"] = 600 * (16 + 68)
    return b


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index; self.lines = []; self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm on host cores
# ------------------------------------------------------------------------------------------------------------------
def _cpu_pair(args):
    """One (frame, partner) pair through the oracle: ORB and lines of both frames, descriptor matching both ways
    (= 2 frames, each with one point match and one line match against its neighbour, or against the map in config 3).
    Frames are synthesised by the caller, outside the timed region."""
    a, b, stages, cfg = args
    from oracle import oracle as orc
    t0 = time.perf_counter()
    res = []
    oc = cfg["orb"]
    orb = orc.OrbOracle(oc["nfeatures"], oc["scale"], oc["nlevels"], oc["ini"], oc["mn"])
    lc = cfg["line"]
    for img in (a, b):
        k, d = orb(img)
        r = [k, d, None, None]
        if "line" in stages and lc:
            ln = orc.LineOracle(lc["nfeatures"], lc["refine"], lc["lsd_scale"], lc["nlevels"], lc["scale"], lc["extractor"])
            r[2], r[3] = ln(img)
        res.append(r)
    if "match" in stages and cfg["match"] == "prev":
        for q, t in ((0, 1), (1, 0)):
            orc.match_ratio(res[q][1], res[t][1], RATIO, MAX_DIST)
            if "line" in stages and lc and len(res[q][3]) and len(res[t][3]):
                orc.match_ratio(res[q][3], res[t][3], RATIO, MAX_DIST)
    elif "match" in stages and cfg["match"] == "map":
        mp_ = _map_descriptors(cfg["n_map"])
        for q in (0, 1):
            orc.match_ratio(res[q][1], mp_, RATIO, MAX_DIST)
    return time.perf_counter() - t0


def _map_descriptors(n):
    return np.random.default_rng(777).integers(0, 256, size=(n, 32), dtype=np.uint8)


def _pair_frames(args):
    seed, h, w = args
    from sdpl_slam_b200 import synth
    f = synth.frame(seed, h, w)
    return f, synth.partner_from(f)


def cpu_baseline(stages, cfg, n_pairs=12):
    """Single host core, bounded sample (n_pairs pairs = 2*n_pairs frames)."""
    h, w = cfg["h"], cfg["w"]
    n_pairs = max(2, int(round(n_pairs * (375 * 1242) / float(h * w))))
    _cpu_pair(_pair_frames((10_000, h, w)) + (stages, cfg))  # warm-up (page-in)
    t = sum(_cpu_pair(_pair_frames((10_001 + i, h, w)) + (stages, cfg)) for i in range(n_pairs))
    return {"value": 2 * n_pairs / t, "unit": UNIT, "cores": 1, "kind": "port", "ms_per_frame": 1000.0 * t / (2 * n_pairs),
            "sample": "%d frames (%d seeded pairs) of the same workload through oracle/liboracle.so, one thread; "
                      "frame synthesis excluded" % (2 * n_pairs, n_pairs)}


def run_reference(args, stages, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    h, w = cfg["h"], cfg["w"]
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    pairs_per_step = max(cores, 2 * ((cores + 1) // 2))
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        # frames are synthesised (in parallel) before the timed region and handed to the workers
        gen = lambda base, n: [f + (stages, cfg) for f in pool.map(_pair_frames, [(base + i, h, w) for i in range(n)])]
        pool.map(_cpu_pair, gen(20_000, cores))                                      # warm-up of every worker
        for w_ in range(args.warmup):
            pool.map(_cpu_pair, gen(30_000 + w_ * pairs_per_step, pairs_per_step), chunksize=1)
        work = [gen(40_000 + s_ * pairs_per_step, pairs_per_step) for s_ in range(args.steps)]
        t0 = time.perf_counter()
        for s_ in range(args.steps):
            pool.map(_cpu_pair, work[s_], chunksize=1)
        dt = time.perf_counter() - t0
    frames = 2 * pairs_per_step * args.steps
    val = frames / dt
    sample = ("each step = %d frames (%d pairs) of the workload, one pair per worker process, %d processes; wall clock "
              "excludes frame synthesis, includes handing the frames to the workers" % (2 * pairs_per_step, pairs_per_step, cores))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": _config(stages, cfg, 2 * pairs_per_step),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def verify_frames(res, himgs, k, cfg):
    """Output check of the timed product path: K frames spread over the last collected batch, every array the
    C ABI returned for them against the oracle on the same frames.  Integers, coordinates, responses, ORB descriptors and
    match indices / distances must be equal; orientations and line end points within 1e-3 (north_star); LBD descriptors equal
    wherever the key line's floats are equal."""
    from oracle import oracle as orc
    F = himgs.shape[0]
    k = max(0, min(int(k), F - 1))
    if k == 0:
        return None
    picks = sorted(set(1 + (j * (F - 1)) // k for j in range(k)))
    oc, lc = cfg["orb"], cfg["line"]
    orb = orc.OrbOracle(oc["nfeatures"], oc["scale"], oc["nlevels"], oc["ini"], oc["mn"])
    line = orc.LineOracle(lc["nfeatures"], lc["refine"], lc["lsd_scale"], lc["nlevels"], lc["scale"], lc["extractor"])
    st = res["stats"]
    fails, lbd_rows, lbd_equal, line_float_equal, nlines = [], 0, 0, 0, 0

    def bad(f, what):
        fails.append("frame %d: %s" % (f, what))

    for f in picks:
        ok_, od = orb(himgs[f]); pk_, pd = orb(himgs[f - 1])
        n = int(st["n_kp"][f])
        gk, gd = res["kps"][f, :n], res["desc"][f, :n]
        if n != len(ok_):
            bad(f, "keypoint count %d vs %d" % (n, len(ok_))); continue
        for name in ("x", "y", "size", "response", "octave", "class_id"):
            if not (gk[name] == ok_[name]).all():
                bad(f, "keypoint field " + name)
        if n and np.abs(gk["angle"] - ok_["angle"]).max() > 1e-3:
            bad(f, "keypoint angle")
        if not (gd == od).all():
            bad(f, "ORB descriptors")
        want, _ = orc.match_ratio(od, pd, RATIO, MAX_DIST)
        gm = res["pt_matches"][f, :n]
        if not ((gm["train"] == want["train"]).all() and (gm["distance"][want["train"] >= 0] == want["distance"][want["train"] >= 0]).all()):
            bad(f, "point matches")
        if int(st["n_pt_matches"][f]) != int((want["train"] >= 0).sum()):
            bad(f, "point match count")
        ol, old = line(himgs[f]); pl, pld = line(himgs[f - 1])
        m = int(st["n_lines"][f])
        gl, gld = res["kls"][f, :m], res["ldesc"][f, :m]
        if m != len(ol):
            bad(f, "line count %d vs %d" % (m, len(ol))); continue
        for name in ("class_id", "octave", "num_pixels"):
            if not (gl[name] == ol[name]).all():
                bad(f, "keyline field " + name)
        same = np.ones(m, bool)
        for name in ("sx", "sy", "ex", "ey", "sx_oct", "sy_oct", "ex_oct", "ey_oct", "pt_x", "pt_y", "length", "response", "size", "angle"):
            tol = 1e-3 if name not in ("response", "size", "angle") else (2e-5 if name == "angle" else None)
            if tol is not None and m and np.abs(gl[name] - ol[name]).max() > tol:
                bad(f, "keyline field " + name)
            same &= gl[name] == ol[name]
        if not (gld[same] == old[same]).all():
            bad(f, "LBD descriptors")
        nlines += m; line_float_equal += int(same.sum()); lbd_rows += m; lbd_equal += int((gld == old).all(1).sum())
        # line matches: the oracle matcher on the descriptors the GPU returned for both frames (equal to the oracle's own
        # wherever the floats are), so that a last-bit difference in one descriptor is not reported twice
        mp_ = int(st["n_lines"][f - 1])
        wantl, _ = orc.match_ratio(gld, res["ldesc"][f - 1, :mp_], RATIO, MAX_DIST) if m and mp_ else (np.zeros(0, orc.DM_DTYPE), 0)
        glm = res["ln_matches"][f, :m]
        if m and mp_ and not (glm["train"] == wantl["train"]).all():
            bad(f, "line matches")
    return {"frames": len(picks), "ok": not fails, "picked": picks, "failures": fails[:8],
            "keylines_checked": nlines, "keylines_float_identical": line_float_equal, "lbd_rows_identical": lbd_equal,
            "what": "ORB keypoints / descriptors, keylines / LBD, ratio-filtered point and line matches of the collected batch vs "
                    "oracle/liboracle.so on the same frames"}


def _config(stages, cfg, frames_per_gpu, extra=None):
    c = {"workload": cfg["workload"], "stages": "+".join(stages), "frames_per_gpu_per_step": frames_per_gpu, "width": cfg["w"],
         "height": cfg["h"],
         "l2": "per-step working set (pyramids, score/blur planes, gradient maps) exceeds the 126 MB L2 many times; "
               "inputs are re-read from HBM every step"}
    if extra:
        c.update(extra)
    return c


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class ResidentArm:
    """The device-resident pipeline over one shard: E = lead + F frames resident in HBM (lead = 1 when the shard has a halo
    frame in front), ORB || lines on two streams, F matching problems (frame t against frame t-1) on two more, per-frame
    statistics + result digests."""

    def __init__(self, fe, torch, dev, local, cfg, stages, F, lead, handles=None, streams=None):
        self.fe, self.torch, self.dev, self.cfg, self.F, self.lead = fe, torch, dev, cfg, F, lead
        oc, lc = cfg["orb"], cfg["line"]
        self.use_line = "line" in stages and lc is not None
        self.use_match = "match" in stages and cfg["match"] is not None
        self.map_mode = cfg["match"] == "map"
        if handles:
            self.orb, self.mat, self.line, self.lmat = handles
        else:
            self.orb = fe.ORBextractor(oc["nfeatures"], oc["scale"], oc["nlevels"], oc["ini"], oc["mn"], device=local)
            self.mat = fe.BinaryDescriptorMatcher(device=local)
            self.line = self.lmat = None
            if self.use_line:
                self.line = fe.Lineextractor(lc["nfeatures"], lc["refine"], lc["lsd_scale"], lc["nlevels"], lc["scale"], lc["extractor"], device=local)
                self.lmat = fe.BinaryDescriptorMatcher(device=local)
        self.cap = cap = self.orb.max_keypoints()
        u8, i32 = torch.uint8, torch.int32
        z = lambda *shape, dt=u8: torch.zeros(shape, dtype=dt, device=dev)
        # descriptor blocks have F+1 slots: slot k = global frame (first owned - 1 + k); slot 0 is the halo frame (or empty)
        self.d_kps = z(F + 1, cap, 28); self.d_desc = z(F + 1, cap, 32); self.d_nkp = z(F + 1, dt=i32)
        self.d_best = z(F, cap, 16); self.d_second = z(F, cap, 16); self.d_out = z(F, cap, 16); self.d_nacc = z(F, dt=i32)
        if self.use_line:
            self.d_kls = z(F + 1, LINE_CAP, 68); self.d_ldesc = z(F + 1, LINE_CAP, 32); self.d_nkl = z(F + 1, dt=i32)
            self.d_lbest = z(F, LINE_CAP, 16); self.d_lsecond = z(F, LINE_CAP, 16); self.d_lout = z(F, LINE_CAP, 16); self.d_lnacc = z(F, dt=i32)
        if self.map_mode:
            self.d_map = torch.from_numpy(_map_descriptors(cfg["n_map"])).to(dev)
            self.d_nmap = torch.full((F,), cfg["n_map"], dtype=i32, device=dev)
        self.stats = z(F, 8, dt=i32)          # n_kp, n_lines, n_pt_matches, n_ln_matches, 2 x 64-bit digests (points, lines)
        self.dig = z(2, F, dt=torch.int64)     # [0]: points, [1]: lines
        if streams:          # shared with another arm (as the two slots of sdpl_frontend share its streams): only the buffers are the arm's own
            self.s_line, self.s_orb, self.s_match, self.s_lmatch = streams
            self.s_tail = torch.cuda.Stream(device=dev, priority=0)
        else:
            self.s_line, self.s_orb, self.s_match, self.s_lmatch, self.s_tail = (torch.cuda.Stream(device=dev, priority=0) for _ in range(5))
        self.launches = 0
        self.bind()

    def bind(self):
        self.orb.set_stream(self.s_orb.cuda_stream); self.mat.set_stream(self.s_match.cuda_stream)
        if self.use_line:
            self.line.set_stream(self.s_line.cuda_stream); self.lmat.set_stream(self.s_lmatch.cuda_stream)

    def handles(self):
        return self.orb, self.mat, self.line, self.lmat

    def streams(self):
        return self.s_line, self.s_orb, self.s_match, self.s_lmatch

    def pair_counts(self):
        """(sum over the F problems of nq * nt) for the point and the line matcher of the last step (host read)."""
        n = self.d_nkp.cpu().numpy().astype(np.int64)
        pt = int((n[1:] * (self.cfg["n_map"] if self.map_mode else n[:-1])).sum())
        ln = 0
        if self.use_line:
            m = self.d_nkl.cpu().numpy().astype(np.int64)
            ln = int((m[1:] * m[:-1]).sum())
        return pt, ln

    def _match(self, m, d, n, rows, best, second, out, nacc):
        F, fs = self.F, rows * 32
        if self.map_mode:
            m.knn2_batch_dev(d.data_ptr() + fs, n.data_ptr() + 4, fs, self.d_map.data_ptr(), self.d_nmap.data_ptr(), 0, F, rows,
                             self.cfg["n_map"], best.data_ptr(), second.data_ptr(), False)
        else:
            m.knn2_batch_dev(d.data_ptr() + fs, n.data_ptr() + 4, fs, d.data_ptr(), n.data_ptr(), fs, F, rows, rows, best.data_ptr(),
                             second.data_ptr(), False)
        self.launches += m.last_launches()
        m.ratio_batch_dev(best.data_ptr(), second.data_ptr(), n.data_ptr() + 4, F, rows, RATIO, MAX_DIST, out.data_ptr(), nacc.data_ptr(), False)
        self.launches += m.last_launches()

    def step(self, d_imgs):
        torch, fe, F, lead, cap = self.torch, self.fe, self.F, self.lead, self.cap
        W, H = self.cfg["w"], self.cfg["h"]
        E = F + lead
        k0 = 1 - lead                       # first slot the extraction writes
        # Everything is enqueued on the arm's own streams behind the caller's current position, and nothing makes the caller's stream
        # wait: two arms used alternately overlap the tail of one step (matching, statistics) with the extraction of the next.
        # The arm's buffers are reused every time: its streams first wait for its previous statistics pass (s_tail)
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(main)
        for s in (self.s_orb, self.s_line, self.s_match, self.s_lmatch):
            s.wait_stream(self.s_tail)
        if self.use_line:
            self.s_line.wait_event(ev)
            self.line.extract_batch_dev(d_imgs.data_ptr(), E, W, H, self.d_kls.data_ptr() + k0 * LINE_CAP * 68,
                                        self.d_ldesc.data_ptr() + k0 * LINE_CAP * 32, LINE_CAP, self.d_nkl.data_ptr() + 4 * k0)
            self.launches += self.line.last_launches()
        self.s_orb.wait_event(ev)
        self.orb.extract_batch_dev(d_imgs.data_ptr(), E, W, H, self.d_kps.data_ptr() + k0 * cap * 28, self.d_desc.data_ptr() + k0 * cap * 32, cap,
                                   self.d_nkp.data_ptr() + 4 * k0)
        self.launches += self.orb.last_launches()
        if self.use_match:
            self.s_match.wait_stream(self.s_orb)
            self._match(self.mat, self.d_desc, self.d_nkp, cap, self.d_best, self.d_second, self.d_out, self.d_nacc)
            if self.use_line and not self.map_mode:
                self.s_lmatch.wait_stream(self.s_line)
                self._match(self.lmat, self.d_ldesc, self.d_nkl, LINE_CAP, self.d_lbest, self.d_lsecond, self.d_lout, self.d_lnacc)
        tail = self.s_tail
        for s in (self.s_orb, self.s_line, self.s_match, self.s_lmatch):
            tail.wait_stream(s)
        with torch.cuda.stream(tail):
            self._stats(tail)
        return self.stats

    def _stats(self, tail):
        """per-frame statistics {n_kp, n_lines, n_point_matches, n_line_matches} + digests of the frame's result rows (on s_tail)"""
        torch, fe, F, cap = self.torch, self.fe, self.F, self.cap
        st, ms = self.stats, tail.cuda_stream
        st[:, 0].copy_(self.d_nkp[1:]); st[:, 2].copy_(self.d_nacc)
        self.dig.zero_()
        dg, dgl = self.dig[0].data_ptr(), self.dig[1].data_ptr()
        n1 = self.d_nkp.data_ptr() + 4
        fe.rows_digest_dev(self.d_kps.data_ptr() + cap * 28, 28, cap * 28, n1, F, cap, 1, dg, ms)
        fe.rows_digest_dev(self.d_desc.data_ptr() + cap * 32, 32, cap * 32, n1, F, cap, 2, dg, ms)
        self.launches += 2
        if self.use_match:
            fe.rows_digest_dev(self.d_out.data_ptr(), 16, cap * 16, n1, F, cap, 3, dg, ms); self.launches += 1
        if self.use_line:
            st[:, 1].copy_(self.d_nkl[1:]); st[:, 3].copy_(self.d_lnacc)
            l1 = self.d_nkl.data_ptr() + 4
            fe.rows_digest_dev(self.d_kls.data_ptr() + LINE_CAP * 68, 68, LINE_CAP * 68, l1, F, LINE_CAP, 4, dgl, ms)
            fe.rows_digest_dev(self.d_ldesc.data_ptr() + LINE_CAP * 32, 32, LINE_CAP * 32, l1, F, LINE_CAP, 5, dgl, ms)
            self.launches += 2
            if self.use_match and not self.map_mode:
                fe.rows_digest_dev(self.d_lout.data_ptr(), 16, LINE_CAP * 16, l1, F, LINE_CAP, 6, dgl, ms); self.launches += 1
        st[:, 4:6].copy_(self.dig[0].view(torch.int32).view(F, 2)); st[:, 6:8].copy_(self.dig[1].view(torch.int32).view(F, 2))

    def wait(self):
        """the caller's current stream waits for the arm's last step"""
        self.torch.cuda.current_stream().wait_stream(self.s_tail)


def stage_table(arm, d_imgs, cfg, peak, facts, reps=3):
    """Per-stage kernel times, each pipeline alone on the device (no cross-stream overlap): CUDA events on the handle's stream."""
    torch, F, lead, cap = arm.torch, arm.F, arm.lead, arm.cap
    W, H = cfg["w"], cfg["h"]
    E, k0 = F + lead, 1 - lead
    stage_ms, stage_launch = {}, {}
    hs = [h for h in arm.handles() if h is not None]
    for h in hs:
        h.set_profiling(True)      # stage marks on the handle's stream (and no internal stream fork: stages run one after another)
    calls = [(arm.orb, "", lambda: arm.orb.extract_batch_dev(d_imgs.data_ptr(), E, W, H, arm.d_kps.data_ptr() + k0 * cap * 28,
                                                             arm.d_desc.data_ptr() + k0 * cap * 32, cap, arm.d_nkp.data_ptr() + 4 * k0))]
    if arm.use_line:
        calls.append((arm.line, "", lambda: arm.line.extract_batch_dev(d_imgs.data_ptr(), E, W, H, arm.d_kls.data_ptr() + k0 * LINE_CAP * 68,
                                                                       arm.d_ldesc.data_ptr() + k0 * LINE_CAP * 32, LINE_CAP,
                                                                       arm.d_nkl.data_ptr() + 4 * k0)))
    if arm.use_match:
        calls.append((arm.mat, "", lambda: arm._match(arm.mat, arm.d_desc, arm.d_nkp, cap, arm.d_best, arm.d_second, arm.d_out, arm.d_nacc)))
        if arm.use_line and not arm.map_mode:
            calls.append((arm.lmat, "line_", lambda: arm._match(arm.lmat, arm.d_ldesc, arm.d_nkl, LINE_CAP, arm.d_lbest, arm.d_lsecond,
                                                                arm.d_lout, arm.d_lnacc)))
    for hdl, prefix, fn in calls:
        for _ in range(reps):
            torch.cuda.synchronize()
            fn()
            for name, t_ms, nl in hdl.stage_times():
                stage_ms[prefix + name] = stage_ms.get(prefix + name, 0.0) + t_ms / reps
                stage_launch[prefix + name] = nl
    torch.cuda.synchronize()
    for h in hs:
        h.set_profiling(False)
    alg = algorithmic_bytes(cfg)
    pt_pairs, ln_pairs = arm.pair_counts() if arm.use_match else (0, 0)
    rows = []
    for name, t_ms in stage_ms.items():
        nbytes = alg.get(name[5:] if name.startswith("line_") else name, 0) * E
        gbs = nbytes / (t_ms * 1e-3) / 1e9 if t_ms > 0 else 0.0
        row = {"stage": name, "ms": round(t_ms, 4), "launches": stage_launch.get(name, 0), "alg_bytes": int(nbytes),
               "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
        if name in ("match_partial", "line_match_partial") and t_ms > 0:
            # the matcher is popc-issue bound, not HBM bound: its roofline is pairs/s against the popc issue rate
            pairs = pt_pairs if name == "match_partial" else ln_pairs
            pps = pairs / (t_ms * 1e-3)
            row.update({"bound": "popc", "pairs": pairs, "pairs_per_s": pps, "popc_per_pair": POPC_PER_PAIR,
                        "popc_frac": round(pps * POPC_PER_PAIR / POPC_PER_S, 4)})
        f = facts.get(name)
        if f:
            row["ncu"] = f
        rows.append(row)
    return rows


def post_stage_rows(fe, torch, dev, local, cfg, arm, peak, F):
    """Frame post-processing on the device (src/Frame.cc:349-389, 482-604, 728-809, 910-925) over the F frames
    of the step, on the key points / key lines the extractors left in HBM: stage times by CUDA events on the handle's stream.
    Not part of the headline step (the metric is ORB + LSD/LBD + match); reported beside it."""
    from sdpl_slam_b200 import synth
    H, W = cfg["h"], cfg["w"]
    nsc = 8
    planes = [synth.scene_planes(s, H, W) for s in range(nsc)]
    rep = (F + nsc - 1) // nsc
    dm = torch.from_numpy(np.stack([p[0] for p in planes])).to(dev).repeat(rep, 1, 1)[:F].contiguous()
    dd = torch.from_numpy(np.stack([p[1] for p in planes])).to(dev).repeat(rep, 1, 1)[:F].contiguous()
    df = torch.from_numpy(np.stack([p[2] for p in planes])).to(dev).repeat(rep, 1, 1, 1)[:F].contiguous()
    post = fe.FramePost(device=local); post.set_profiling(True)
    step = 4
    cap = ((H + step - 1) // step) * ((W + step - 1) // step)
    u8, i32, f32 = torch.uint8, torch.int32, torch.float32
    z = lambda *shape, dt=u8: torch.zeros(shape, dtype=dt, device=dev)
    keys, corres, fn, dep, lab, n = z(F, cap, 28), z(F, cap, 28), z(F, cap, 2, dt=f32), z(F, cap, dt=f32), z(F, cap, dt=i32), z(F, dt=i32)
    KC, LC = arm.cap, LINE_CAP
    rows = []

    def timed(name, fn_, reps=3):
        t = 0.0
        for _ in range(reps):
            torch.cuda.synchronize(); fn_()
            t += sum(ms for _, ms, _ in post.stage_times()) / reps
        return t

    t = timed("object_sampling", lambda: post.sample_objects_dev(dm.data_ptr(), dd.data_ptr(), df.data_ptr(), F, W, H, step, 25.0, keys.data_ptr(),
                                                                 corres.data_ptr(), fn.data_ptr(), dep.data_ptr(), lab.data_ptr(), cap, n.data_ptr()))
    kept = int(n.sum().item())
    samp = dm[:, ::step, ::step]
    nonzero = int((samp != 0).sum().item()); nsamp = samp.numel()
    addressed = nsamp * 4 + nonzero * 12 + kept * (72 + 12 + 4)          # mask; depth + flow where the mask is set; records (+ re-read, bitmap)
    sectors = F * ((H + step - 1) // step) * W * 4 + nonzero * 64 + kept * (72 + 64)   # 32-byte sectors: every sampled mask row in full, one sector
    rows.append({"stage": "object_sampling", "ms": round(t, 4), "launches": 3, "alg_bytes": int(addressed), "gbs": round(addressed / t / 1e6, 1),
                 "frac": round(addressed / t / 1e6 / peak, 4), "sector_bytes": int(sectors), "sector_gbs": round(sectors / t / 1e6, 1),
                 "frac_sectors": round(sectors / t / 1e6 / peak, 4), "samples_kept_per_frame": kept / F,
                 "note": "every 4th pixel of every 4th row: a 32-byte sector holds two mask samples, so the DRAM traffic of the compulsory "
                         "sectors (sector_bytes) is 4x the addressed bytes"})
    if arm.use_line:
        fk, fidx, fnn = z(F, LC, 68), z(F, LC, dt=i32), z(F, dt=i32)
        kl1 = arm.d_kls.data_ptr() + LC * 68; nl1 = arm.d_nkl.data_ptr() + 4
        t = timed("line_filters", lambda: post.filter_lines_dev(dm.data_ptr(), dd.data_ptr(), F, W, H, kl1, nl1, LC, fk.data_ptr(), fidx.data_ptr(),
                                                                fnn.data_ptr()))
        rows.append({"stage": "line_filters", "ms": round(t, 4), "launches": 1})
        obj, nobj, stat, cor = z(F, LC, 68), z(F, dt=i32), z(F, LC, 68), z(F, LC, 68)
        lfn, inf, sdep, sidx, nst = z(F, LC, 4, dt=f32), z(F, LC, 3, dt=torch.float64), z(F, LC, 2, dt=f32), z(F, LC, dt=i32), z(F, dt=i32)
        t = timed("line_corres", lambda: post.line_corres_dev(dm.data_ptr(), dd.data_ptr(), df.data_ptr(), F, W, H, fk.data_ptr(), fnn.data_ptr(), LC, 40.0,
                                                              obj.data_ptr(), nobj.data_ptr(), stat.data_ptr(), cor.data_ptr(), lfn.data_ptr(),
                                                              inf.data_ptr(), sdep.data_ptr(), sidx.data_ptr(), nst.data_ptr()))
        rows.append({"stage": "line_corres", "ms": round(t, 4), "launches": 1, "lines_kept_per_frame": float(nst.float().mean().item())})
    kp1 = arm.d_kps.data_ptr() + KC * 28; nk1 = arm.d_nkp.data_ptr() + 4
    pst, pcor, pfn, pdep, pidx, npt = z(F, KC, 28), z(F, KC, 28), z(F, KC, 2, dt=f32), z(F, KC, dt=f32), z(F, KC, dt=i32), z(F, dt=i32)
    t = timed("point_corres", lambda: post.point_corres_dev(dm.data_ptr(), dd.data_ptr(), df.data_ptr(), F, W, H, kp1, nk1, KC, 40.0, pst.data_ptr(),
                                                            pcor.data_ptr(), pfn.data_ptr(), pdep.data_ptr(), pidx.data_ptr(), npt.data_ptr()))
    rows.append({"stage": "point_corres", "ms": round(t, 4), "launches": 1, "points_kept_per_frame": float(npt.float().mean().item())})
    cs, items = z(F, 64 * 48 + 1, dt=i32), z(F, KC, dt=i32)
    t = timed("grid", lambda: post.grid_dev(F, W, H, kp1, nk1, KC, cs.data_ptr(), items.data_ptr(), 64, 48))
    rows.append({"stage": "grid", "ms": round(t, 4), "launches": 1})
    torch.cuda.synchronize()
    return {"what": "Frame::Frame post-processing on the device over the step's frames (src/Frame.cc:349-389, 482-604, 728-809, 910-925), "
                    "%d distinct mask / depth / flow plane sets; not inside the timed step" % nsc,
            "total_ms": round(sum(r["ms"] for r in rows), 4), "stages": rows}


def run_gpu(args, stages, cfg):
    import torch
    import torch.distributed as dist
    from sdpl_slam_b200 import frontend as fe, shard, synth

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the front-end has no CPU fallback)")
    H, W = cfg["h"], cfg["w"]
    # ---- the batch and this rank's shard of it (synthesised before CUDA / NCCL start: the pool forks) ----
    if args.scaling == "strong":
        total = args.total_frames
    else:
        total = world * (args.frames if args.frames > 0 else cfg["frames"])
    h0, s0, e0_ = shard.shard_with_halo(total, rank, world)
    F, lead = e0_ - s0, s0 - h0
    host = synth.sequence(h0, e0_, H, W, workers=min(16, max(1, (os.cpu_count() or 2) // max(1, world))))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    pinned = torch.from_numpy(host).pin_memory()
    d_imgs = pinned.to(dev)
    # two sets of output buffers used alternately, so that step k + 1's extraction does not wait for step k's matching and statistics.
    # --pipeline 1 (default): ONE set of handles and streams, as the two slots of sdpl_frontend_submit / collect (the extractor streams run
    # step after step, the matchers and statistics of step k beside the extraction of step k + 1); --pipeline 2: two complete sets of
    # handles (two region-growing kernels may then run side by side: measured slower)
    n_arms = 2 if (args.pipeline and F <= 1024) else 1
    arms = [ResidentArm(fe, torch, dev, local, cfg, stages, F, lead)]
    if n_arms == 2:
        arms.append(ResidentArm(fe, torch, dev, local, cfg, stages, F, lead) if args.pipeline >= 2 else
                    ResidentArm(fe, torch, dev, local, cfg, stages, F, lead, handles=arms[0].handles(), streams=arms[0].streams()))
    gathers = [shard.StatsGather(total, 8, torch.int32, dev) for _ in range(n_arms)]
    arm = arms[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_step(k):
        a, g = arms[k % n_arms], gathers[k % n_arms]
        st_ = a.step(d_imgs)
        with torch.cuda.stream(a.s_tail):
            g.start(st_)                    # NCCL all-gather of 32 B per frame: the only collective of the path (asynchronous)
        return a

    def drain():
        for a, g in zip(arms, gathers):
            a.wait()
            g.table()

    # ---- device-resident throughput (`value`) ----
    nwarm = max(args.warmup, 3)
    for k in range(nwarm):
        run_step(k)
    drain()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for a in arms:
        a.launches = 0
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 2)]
    evs[0].record()
    for k in range(args.steps):
        a = run_step(k)
        evs[k + 1].record(a.s_tail)
    drain()                                 # the current stream waits for both arms and their last gathers
    table = gathers[(args.steps - 1) % n_arms].table()
    evs[args.steps + 1].record()
    barrier()
    ms = evs[0].elapsed_time(evs[args.steps + 1])
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    n_launch = sum(a.launches for a in arms)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    mine = torch.tensor([float(np.min(step_ms)), float(np.median(step_ms)), float(np.max(step_ms)), ms], dtype=torch.float64, device=dev)
    per_rank = mine.unsqueeze(0)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allr = torch.zeros((world, 4), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine.unsqueeze(0))
        per_rank = allr
    ms_max = float(t.item())
    value = total * args.steps / (ms_max / 1000.0)
    table = table.cpu().numpy()
    st = table[:, :4]
    pr = per_rank.cpu().numpy()
    rank_ms = {"per_rank_step_ms_median": [round(float(x), 3) for x in pr[:, 1]],
               "own_step_ms": {"min": round(float(pr[:, 0].min()), 3), "median": round(float(np.median(pr[:, 1])), 3), "max": round(float(pr[:, 2].max()), 3)},
               "per_rank_total_ms": [round(float(x), 3) for x in pr[:, 3]], "timed_ms_max_over_ranks": round(ms_max, 3),
               "arms": n_arms,
               "note": "a rank's own step = interval between the completions (CUDA events on the statistics stream) of consecutive steps; "
                       "with two arms consecutive steps overlap (extraction of step k+1 behind matching / statistics of step k), so the first "
                       "interval is longer than the rest; the asynchronous gather is not inside; region growing is data dependent, so ranks "
                       "differ; timed_ms includes the wait for the last gather"}

    # ---- the sharded table must equal what ONE GPU computes for the same frames: rank 0 recomputes the first two frames of every
    #      other rank's shard (with their predecessor) and compares counts and 64-bit result digests with the gathered table ----
    shard_check = None
    if world > 1 and rank == 0:
        bad, checked = [], 0
        for r in range(1, world):
            hs_, ss_, es_ = shard.shard_with_halo(total, r, world)
            nchk = min(2, es_ - ss_)
            if nchk < 1 or ss_ == hs_:
                continue
            small = ResidentArm(fe, torch, dev, local, cfg, stages, nchk, 1, handles=arm.handles())
            imgs_small = torch.from_numpy(synth.sequence(hs_, ss_ + nchk, H, W)).to(dev)
            got = small.step(imgs_small)
            torch.cuda.synchronize()
            got = got.cpu().numpy()
            for j in range(nchk):
                checked += 1
                if not (got[j] == table[ss_ + j]).all():
                    bad.append({"frame": ss_ + j, "rank": r, "one_gpu": got[j].tolist(), "sharded": table[ss_ + j].tolist()})
            del small
        arm.bind()
        shard_check = {"frames": checked, "ok": not bad, "failures": bad[:4],
                       "what": "per-frame counts + 64-bit digests of keypoints / descriptors / keylines / LBD / filtered matches in the "
                               "NCCL-gathered table vs the same frames recomputed on rank 0 alone"}

    peak, peak_src = _peaks()
    facts = _ncu_facts()
    stage_rows = stage_table(arm, d_imgs, cfg, peak, facts.get("stages", {}))
    post_block = None
    if rank == 0 and not args.no_post and cfg["baseline"] == 1:
        post_block = post_stage_rows(fe, torch, dev, local, cfg, arm, peak, F)
        torch.cuda.empty_cache()
    dom = max(stage_rows, key=lambda r: r["ms"]) if stage_rows else None
    roofline = None
    if dom:
        nl = max(1, dom["launches"])
        nf = (dom.get("ncu") or {})
        tpf = nf.get("dram_bytes_per_frame")
        roofline = {"bound": "hbm", "kernel": dom["stage"], "achieved": dom["gbs"], "peak": peak, "unit": "GB/s", "frac": dom["frac"],
                    "traffic": int(tpf * (F + lead) / nl) if tpf else None, "traffic_source": nf.get("source"), "peak_source": peak_src,
                    "launches_per_step": nl, "avg_launch_ms": dom["ms"] / nl, "alg_bytes_per_launch": dom["alg_bytes"] // nl,
                    "note": "stage time by CUDA events on the handle's stream, pipeline run alone, mean of 3 calls; the dominant kernel is a "
                            "dependent-instruction chain (greedy region growing), bounded by latency and by the CTAs an SM can hold, not by "
                            "bandwidth; traffic = ncu DRAM bytes per frame (profiles/ncu_stage_facts.json) x frames per launch"}

    # ---- end to end through the host-buffer C ABI (`e2e`): pinned host frames in, results back on the host ----
    e2e = verified = latency = None
    full = arm.use_line and arm.use_match and not arm.map_mode
    frame_stats_mean = {"keypoints": float(st[:, 0].mean()), "keylines": float(st[:, 1].mean()),
                        "point_matches": float(st[:, 2].mean()), "line_matches": float(st[:, 3].mean())}
    if not args.no_e2e and full and args.scaling != "strong":
        for a in arms:
            for hdl in a.handles():
                hdl.set_stream(0)
        del arm, a
        arms.clear()
        torch.cuda.empty_cache()
        oc, lc = cfg["orb"], cfg["line"]
        front = fe.FrontEnd(oc["nfeatures"], oc["scale"], oc["nlevels"], oc["ini"], oc["mn"], lc["nfeatures"],
                            lc["refine"], lc["lsd_scale"], lc["nlevels"], lc["scale"], RATIO, MAX_DIST, device=local)
        if lc["extractor"]:
            front.set_line_extractor(lc["extractor"])
        himgs = pinned.numpy()[lead:]            # the F owned frames; frame 0 of a batch is matched against the previous batch's last
        e2e_steps = max(1, min(args.steps, 5))
        res = front.process(himgs)
        res = front.process(himgs)
        barrier()
        # pipelined use of the same API: the next batch is submitted before the current one is collected, so its upload and
        # the previous batch's download overlap the kernels; all e2e_steps batches run entirely inside the timed region
        t0 = time.perf_counter()
        front.submit(himgs)
        for _ in range(e2e_steps - 1):
            front.submit(himgs)
            res = front.collect()
        res = front.collect()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        est = res["stats"]
        d2h = int(est["n_kp"].sum()) * (28 + 32 + 16) + int(est["n_lines"].sum()) * (68 + 32 + 16) + 16 * F
        e2e = {"value": world * F * e2e_steps / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(himgs.nbytes),
               "d2h_bytes_per_step": d2h, "steps": e2e_steps, "gpu_launches_per_step": front.last_launches(),
               "api": "FrontEnd.submit / collect (sdpl_frontend_submit / sdpl_frontend_collect; process = both): pinned host frames "
                      "in (one H2D per batch), keypoints + ORB descriptors + keylines + LBD descriptors + ratio-filtered matches + "
                      "per-frame counts back on the host; batch k+1 is submitted before batch k is collected",
               "frame_stats_mean": {"keypoints": float(est["n_kp"].mean()), "keylines": float(est["n_lines"].mean()),
                                    "point_matches": float(est["n_pt_matches"].mean()), "line_matches": float(est["n_ln_matches"].mean())}}
        if rank == 0 and args.verify > 0:
            verified = verify_frames(res, himgs, args.verify, cfg)
        del front
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_latency:
        latency = latency_block(fe, cfg, stages, pinned.numpy()[lead:lead + 16], local)

    if rank == 0:
        cpu = None if args.no_cpu else cpu_baseline(stages, cfg)
        if latency and cpu:
            latency["cpu_one_core_ms_per_frame"] = round(cpu["ms_per_frame"], 2)
            latency["speedup_vs_one_core"] = round(cpu["ms_per_frame"] / latency["ms_per_frame_median"], 2)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": nwarm,
               "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u8",
               "data": "synthetic (seeded noise-texture + rectangles frames, sdpl_slam_b200/synth.py: sequence; %d distinct frames in the "
                       "batch, %d per GPU%s)" % (total, F, " + 1 halo frame on ranks > 0" if world > 1 else ""),
               "config": _config(stages, cfg, F, {"total_frames_per_step": total, "sharding": "shard.shard_with_halo (contiguous blocks)"}),
               "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch, "roofline": roofline,
               "cpu_baseline": cpu, "verified": verified, "latency": latency, "ranks": rank_ms, "sharding_check": shard_check,
               "stages": stage_rows, "post": post_block, "frame_stats_mean": frame_stats_mean}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def latency_block(fe, cfg, stages, frames, local, n=50):
    """The reference's own call pattern (src/Frame.cc:927-949): ONE frame per call through the host-buffer C ABI --
    sdpl_orb_extract, sdpl_line_extract, sdpl_match_ratio (points and lines against the previous frame) -- with pageable numpy
    buffers, synchronous, nothing overlapped.  Median over n frames after 5 warm-up frames, milliseconds."""
    oc, lc = cfg["orb"], cfg["line"]
    orb = fe.ORBextractor(oc["nfeatures"], oc["scale"], oc["nlevels"], oc["ini"], oc["mn"], device=local)
    use_line = "line" in stages and lc is not None
    line = fe.Lineextractor(lc["nfeatures"], lc["refine"], lc["lsd_scale"], lc["nlevels"], lc["scale"], lc["extractor"], device=local) if use_line else None
    m = fe.BinaryDescriptorMatcher(device=local)
    mapd = _map_descriptors(cfg["n_map"]) if cfg["match"] == "map" else None
    frames = [np.array(f) for f in frames]                     # pageable copies
    t_orb, t_line, t_match, t_all = [], [], [], []
    pd = pld = None
    for i in range(n + 5):
        img = frames[i % len(frames)]
        t0 = time.perf_counter()
        k, d = orb(img)
        t1 = time.perf_counter()
        dl = None
        if line is not None:
            kl, dl = line(img, capacity=4096)
        t2 = time.perf_counter()
        if "match" in stages and cfg["match"]:
            if mapd is not None:
                m.ratioMatch(d, mapd, RATIO, MAX_DIST)
            elif pd is not None:
                m.ratioMatch(d, pd, RATIO, MAX_DIST)
                if dl is not None and pld is not None and len(dl) and len(pld):
                    m.ratioMatch(dl, pld, RATIO, MAX_DIST)
        t3 = time.perf_counter()
        pd, pld = d, dl
        if i >= 5:
            t_orb.append(t1 - t0); t_line.append(t2 - t1); t_match.append(t3 - t2); t_all.append(t3 - t0)
    med = lambda a: round(1000.0 * float(np.median(a)), 3)
    return {"batch": 1, "frames": n, "ms_per_frame_median": med(t_all), "ms_per_frame_p90": round(1000.0 * float(np.percentile(t_all, 90)), 3),
            "orb_ms": med(t_orb), "line_ms": med(t_line), "match_ms": med(t_match), "fps": round(1.0 / float(np.median(t_all)), 1),
            "api": "sdpl_orb_extract + sdpl_line_extract + sdpl_match_ratio, one frame per call, pageable host buffers, synchronous"}


def run_sweep(args, stages, cfg):
    """Batch-size sweep through the host-buffer API (FrontEnd.process, nothing pipelined across calls): ms per frame as a
    function of the frames per call."""
    import torch
    from sdpl_slam_b200 import frontend as fe, synth
    H, W = cfg["h"], cfg["w"]
    oc, lc = cfg["orb"], cfg["line"]
    base = synth.sequence(0, 512, H, W, workers=min(16, os.cpu_count() or 1))
    rows = []
    lat = latency_block(fe, cfg, stages, base[:16], 0)
    rows.append({"batch": 1, "api": "sdpl_orb_extract + sdpl_line_extract + sdpl_match_ratio", "ms_per_call": lat["ms_per_frame_median"],
                 "ms_per_frame": lat["ms_per_frame_median"], "fps": lat["fps"]})
    front = fe.FrontEnd(oc["nfeatures"], oc["scale"], oc["nlevels"], oc["ini"], oc["mn"], lc["nfeatures"], lc["refine"], lc["lsd_scale"],
                        lc["nlevels"], lc["scale"], RATIO, MAX_DIST, device=0)
    for B in (1, 8, 64, 512, 2048):
        imgs = torch.from_numpy(np.concatenate([base] * (B // 512)) if B > 512 else base[:B].copy()).pin_memory().numpy()
        reps = max(3, min(30, 256 // B))
        front.process(imgs); front.process(imgs)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); front.process(imgs); ts.append(time.perf_counter() - t0)
        t = float(np.median(ts))
        rows.append({"batch": B, "api": "sdpl_frontend_process (pinned host frames, one call)", "ms_per_call": round(1000 * t, 3),
                     "ms_per_frame": round(1000 * t / B, 4), "fps": round(B / t, 1)})
    print(json.dumps({"metric": METRIC, "unit": UNIT, "sweep": rows, "config": _config(stages, cfg, 0), "data": "synthetic",
                      "note": "host-buffer API, synchronous calls; batch 1 = the reference's one-frame-per-call pattern"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="SURVEY.md 8d numbering; 2 = BASELINE.json configs[1] (the metric's)")
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (weak scaling); default: the config's (512)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --frames per GPU (default); strong: --total-frames in total at every N (BASELINE configs[3] read literally)")
    ap.add_argument("--total-frames", type=int, default=4096)
    ap.add_argument("--stages", default="orb,line,match")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--pipeline", type=int, default=1, help="1: two sets of output buffers on one set of handles / streams, consecutive steps overlap "
                                                              "(when a GPU holds <= 1024 frames per step); 2: two complete sets of handles; 0: steps strictly one after another")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-post", action="store_true", help="skip the Frame post-processing stage block (SURVEY 8f rows 1, 2)")
    ap.add_argument("--sweep", action="store_true", help="batch-size sweep {1, 8, 64, 512, 2048} through the host-buffer API (1 GPU)")
    ap.add_argument("--extractor", type=int, default=0, choices=(0, 1),
                    help="line back-end of Lineextractor: 0 = LSD (BASELINE.json's configuration), 1 = EDLines (SURVEY 8f row 3)")
    ap.add_argument("--verify", type=int, default=8, help="frames of the last end-to-end batch compared with the oracle (0 = off)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.extractor and cfg.get("line"):
        cfg = dict(cfg, line=dict(cfg["line"], extractor=args.extractor),
                   workload=cfg["workload"] + " -- with the EDLines line back-end (extractor = 1) instead of LSD")
    stages = [s for s in args.stages.split(",") if s]
    if cfg["line"] is None:
        stages = [s for s in stages if s != "line"]
    if cfg["match"] is None:
        stages = [s for s in stages if s != "match"]
    if args.impl == "reference":
        run_reference(args, stages, cfg)
    elif args.sweep:
        run_sweep(args, stages, cfg)
    else:
        run_gpu(args, stages, cfg)


if __name__ == "__main__":
    main()
