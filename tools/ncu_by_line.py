"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line: warp instructions executed,
thread instructions, stall samples.  Usage: python tools/ncu_by_line.py dump.csv [file-substring] [top-n]"""
import csv, sys, collections
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])   # (file,line) -> inst, thread inst, samples, text
tot = [0, 0, 0]
i = 0
cur_line = None
while i < len(rows):
    r = rows[i]
    if r and r[0] in ("File Name", "File Path"):
        cur_file = r[1]; i += 1; continue
    if r and r[0] == "Line No":
        hdr = r; i += 1; continue
    if hdr and len(r) >= 10:
        ln, src, addr = r[0], r[1], r[2]
        if ln != "":
            cur_line = (cur_file, int(ln)); agg[cur_line][3] = src
        if addr != "" and cur_line:
            try:
                inst = int(r[hdr.index("Instructions Executed")] or 0); ti = int(r[hdr.index("Thread Instructions Executed")] or 0)
                smp = int(r[hdr.index("# Samples")] or 0)
            except ValueError:
                inst = ti = smp = 0
            a = agg[cur_line]; a[0] += inst; a[1] += ti; a[2] += smp
            tot[0] += inst; tot[1] += ti; tot[2] += smp
    i += 1
print("total warp inst %d, thread inst %d (%.1f thr/inst), samples %d" % (tot[0], tot[1], tot[1] / max(1, tot[0]), tot[2]))
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), a in agg.items():
    b = byfile[f]; b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("%-60s inst %5.1f%% samples %5.1f%%" % (str(f)[-60:], 100 * b[0] / tot[0], 100 * b[2] / max(1, tot[2])))
items = [(k, a) for k, a in agg.items() if want in str(k[0])]
items.sort(key=lambda kv: -kv[1][2])
print("--- top lines by stall samples")
for (f, l), a in items[:topn]:
    print("%s:%d  inst %5.2f%%  thr/inst %5.1f  samples %5.2f%%  | %s" % (str(f).split('/')[-1], l, 100 * a[0] / tot[0], a[1] / max(1, a[0]), 100 * a[2] / max(1, tot[2]), a[3].strip()[:110]))
if len(sys.argv) > 4:
    # buckets: "name:lo-hi,..." over line numbers (single-file kernels)
    print("--- buckets")
    for spec in sys.argv[4].split(","):
        name, rng = spec.split(":"); lo, hi = map(int, rng.split("-"))
        fsub = name.split("@")[1] if "@" in name else ""
        a = [0, 0, 0]
        for (f, l), v in agg.items():
            if lo <= l <= hi and fsub in str(f):
                a[0] += v[0]; a[1] += v[1]; a[2] += v[2]
        print("%-28s inst %5.1f%%  thr/inst %5.1f  samples %5.1f%%" % (name, 100 * a[0] / tot[0], a[1] / max(1, a[0]), 100 * a[2] / max(1, tot[2])))
