"""Turn the files a GPU run left in gpurun_out/ into the tracked summaries under profiles/ (bench JSON lines, ncu launch list,
ncu --set full extracts).  Usage: python tools/summarize_profiles.py"""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']


def full(rep, out, header):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    idx = [h.index(k) for k in WANT if k in h]
    with open(out, 'w', newline='') as f:
        f.write('"# %s"\n' % header)
        w = csv.writer(f)
        w.writerow([h[i] for i in idx]); w.writerow([u[i] for i in idx])
        for v in rows[2:]:
            w.writerow([v[i] for i in idx])


def launches(src, out_csv, out_txt, header):
    shutil.copy(src, out_csv)
    rows = list(csv.reader(l for l in open(src) if not l.startswith('==')))
    h = rows[0]; ki, mi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg, cnt = {}, collections.Counter()
    for x in rows[1:]:
        if len(x) <= mi:
            continue
        name = re.sub(r'\(.*', '', x[ki])[:60]; v = float(x[mi].replace(',', '')); u = x[ui]
        ms = v / 1e6 if u.startswith('n') else (v / 1e3 if u.startswith('u') else v)
        agg[name] = agg.get(name, 0) + ms; cnt[name] += 1
    tot = sum(agg.values())
    out = [header, "per-launch times are serialised and cold-cache: compare shares", ""]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
        out.append(f"{k:62s} {cnt[k]:3d} launches {v:10.2f} ms {100 * v / tot:5.1f}%")
    open(out_txt, 'w').write("\n".join(out) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "final"
    for a, b in (("bench_final.json", "r1_bench_%s.json" % tag), ("bench_ref.json", "r1_bench_reference_arm.json"), ("bench_8gpu.json", "r1_bench_%s_8gpu.json" % tag)):
        if os.path.exists(os.path.join(G, a)):
            shutil.copy(os.path.join(G, a), os.path.join(P, b))
    if os.path.exists(os.path.join(G, "launches_final2.csv")):
        launches(os.path.join(G, "launches_final2.csv"), os.path.join(P, "r1_launches_%s.csv" % tag), os.path.join(P, "r1_launches_%s_summary.txt" % tag),
                 "ncu --metrics gpu__time_duration.sum --clock-control none -c 900 : python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e (512 frames per step, 1 GPU, %s round-1 build)" % tag)
    if os.path.exists(os.path.join(G, "orb_full.ncu-rep")):
        full(os.path.join(G, "orb_full.ncu-rep"), os.path.join(P, "r1_orb_kernels_%s_ncu_full_summary.csv" % tag),
             "ncu --set full --clock-control none, ORB kernels of the %s round-1 build, 128 frames 1242x375: python tools/prof_orb.py 128" % tag)
    if os.path.exists(os.path.join(G, "line_full.ncu-rep")):
        full(os.path.join(G, "line_full.ncu-rep"), os.path.join(P, "r1_lsd_kernels_%s_ncu_full_summary.csv" % tag),
             "ncu --set full --clock-control none, k_lsd_grow_block<4,4> and k_lsd_nfa_rest of the %s round-1 build, 128 frames 1242x375: python tools/prof_grow2.py 128 0" % tag)
