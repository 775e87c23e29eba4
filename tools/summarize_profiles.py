"""Turn what tools/ncu_capture.sh left in gpurun_out/ into the tracked summaries under profiles/:
  <tag>_kernels_ncu_full_summary.csv   selected `ncu --set full` metrics of every product kernel (one pass over F frames)
  <tag>_launches.csv / _summary.txt    launch list of a short bench run (per-kernel time shares)
  ncu_stage_facts.json                 per bench stage: DRAM bytes per frame, issue-slot utilisation ... (read by bench.py)
Usage: python tools/summarize_profiles.py TAG FRAMES [GIT_SHA]"""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
# bench stage -> kernels of the stage
STAGES = {"pyramid": ["k_pyr_base", "k_pyr_resize"], "fast_score": ["k_fast_score"], "cell_nms": ["k_cell_nms", "k_cell_emit"], "quadtree": ["k_quadtree"],
          "blur7": ["k_blur7"], "orient_describe": ["k_orient_describe"], "lsd_pyramid": ["k_line_resize"], "lsd_scale": ["k_lsd_scale"],
          "lsd_gradient": ["k_lsd_grad"], "lsd_sort": ["k_lsd_sort", "k_lsd_sort_scan", "k_lsd_task_rank"], "lsd_grow": ["k_lsd_grow2"],
          "lsd_nfa": ["k_lsd_nfa", "k_lsd_nfa_big", "k_lsd_nfa_rest"], "keylines": ["k_keylines"],
          "lbd_sobel": ["k_lbd_blur_sobel", "k_lbd_pyrdown", "k_lbd_sobel"], "lbd_bands": ["k_lbd"], "match_partial": ["k_match_partial"],
          "match_merge": ["k_match_merge"], "object_sampling": ["k_obj_flags", "k_obj_scan", "k_obj_emit"]}


def num(x):
    try:
        return float(x.replace(',', ''))
    except ValueError:
        return 0.0


def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)


def kname(full):
    m = re.search(r'(k_[a-z0-9_]+)', full)
    return m.group(1) if m else full[:40]


def full_summary(raw_csv, out_csv, header, frames, source):
    rows = list(csv.reader(open(raw_csv)))
    h, u = rows[0], rows[1]
    idx = [h.index(k) for k in WANT if k in h]
    seen, keep = collections.Counter(), []
    for v in rows[2:]:
        n = kname(v[h.index('Kernel Name')])
        if not n.startswith("k_"):
            continue
        keep.append(v)
    with open(out_csv, 'w', newline='') as f:
        f.write('"# %s"\n' % header)
        w = csv.writer(f)
        w.writerow([h[i] for i in idx]); w.writerow([u[i] for i in idx])
        for v in keep:
            w.writerow([v[i] for i in idx])
    # per-stage facts (the capture runs every kernel twice -- second matcher / LBD pass of prof_all.py; average per kernel launch index)
    per = collections.defaultdict(lambda: collections.defaultdict(list))
    ci = {k: h.index(k) for k in WANT if k in h}
    for v in keep:
        n = kname(v[ci['Kernel Name']])
        per[n]['t'].append(num(v[ci['gpu__time_duration.sum']]) * unit_scale(u[ci['gpu__time_duration.sum']]))
        per[n]['rd'].append(num(v[ci['dram__bytes_read.sum']]) * unit_scale(u[ci['dram__bytes_read.sum']]))
        per[n]['wr'].append(num(v[ci['dram__bytes_write.sum']]) * unit_scale(u[ci['dram__bytes_write.sum']]))
        per[n]['iss'].append(num(v[ci['smsp__issue_active.avg.pct_of_peak_sustained_active']]))
        per[n]['thr'].append(num(v[ci['smsp__thread_inst_executed_per_inst_executed.ratio']]))
        per[n]['occ'].append(num(v[ci['sm__warps_active.avg.pct_of_peak_sustained_active']]))
    facts = {}
    for stage, ks in STAGES.items():
        t = rd = wr = 0.0; iss_w = thr_w = occ_w = 0.0; found = False
        for k in ks:
            if k not in per:
                continue
            found = True
            d = per[k]
            # kernels launched once per level / octave appear several times in ONE pass; the capture holds one pass of the pipeline kernels
            # and two of the ones prof_all.py repeats (matcher, LBD, post): normalise those by their repeat count
            rep = 2 if (k.startswith(("k_match", "k_obj", "k_lbd", "k_point", "k_grid")) and len(d['t']) % 2 == 0) else 1
            kt = sum(d['t']) / rep
            t += kt; rd += sum(d['rd']) / rep; wr += sum(d['wr']) / rep
            iss_w += sum(a * b for a, b in zip(d['iss'], d['t'])) / rep
            thr_w += sum(a * b for a, b in zip(d['thr'], d['t'])) / rep
            occ_w += sum(a * b for a, b in zip(d['occ'], d['t'])) / rep
        if found and t > 0:
            facts[stage] = {"dram_bytes_per_frame": (rd + wr) / frames, "dram_read_bytes_per_frame": rd / frames, "dram_write_bytes_per_frame": wr / frames,
                            "issue_active_pct": round(iss_w / t, 1), "threads_per_instruction": round(thr_w / t, 1),
                            "warps_active_pct": round(occ_w / t, 1), "ncu_ms_cold": round(1e3 * t, 4), "frames": frames, "source": source}
    return facts


def launches(src, out_csv, out_txt, header):
    shutil.copy(src, out_csv)
    rows = list(csv.reader(l for l in open(src) if not l.startswith('==')))
    h = rows[0]; ki, mi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg, cnt = {}, collections.Counter()
    for x in rows[1:]:
        if len(x) <= mi:
            continue
        name = re.sub(r'\(.*', '', x[ki])[:60]; v = num(x[mi]); u = x[ui]
        ms = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
        agg[name] = agg.get(name, 0) + ms; cnt[name] += 1
    tot = sum(agg.values())
    out = [header, "per-launch times are serialised and cold-cache: compare shares", ""]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        out.append(f"{k:62s} {cnt[k]:3d} launches {v:10.2f} ms {100 * v / tot:5.1f}%")
    open(out_txt, 'w').write("\n".join(out) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1]; frames = int(sys.argv[2]); sha = sys.argv[3] if len(sys.argv) > 3 else subprocess.run(
        ["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    raw = os.path.join(G, "%s_all_raw.csv" % tag)
    if os.path.exists(raw):
        src = "profiles/%s_kernels_ncu_full_summary.csv (ncu --set full --clock-control none, tools/prof_all.py %d, build %s)" % (tag, frames, sha)
        facts = full_summary(raw, os.path.join(P, "%s_kernels_ncu_full_summary.csv" % tag),
                             "ncu --set full --clock-control none: python tools/prof_all.py %d (one pass of every product kernel over %d frames 1242x375 after a "
                             "warm-up pass; build %s)" % (frames, frames, sha), frames, src)
        json.dump({"build": sha, "frames": frames, "stages": facts}, open(os.path.join(P, "ncu_stage_facts.json"), "w"), indent=1)
    ll = os.path.join(G, "%s_launches.csv" % tag)
    if os.path.exists(ll):
        launches(ll, os.path.join(P, "%s_launches.csv" % tag), os.path.join(P, "%s_launches_summary.txt" % tag),
                 "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : python bench.py --steps 1 --warmup 3 --pipeline 0 --no-cpu --no-e2e "
                 "--no-latency --no-post (512 frames per step, 1 GPU, build %s)" % sha)
