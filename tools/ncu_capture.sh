#!/bin/bash
# Runs ON THE GPU BOX (gpurun): `ncu --set full` over one pass of every product kernel (tools/prof_all.py), keeps the report in /tmp and
# brings back only the raw-metrics CSV (gpurun_out/ is limited to 64 MiB), plus the launch list of a short bench run.
# Usage: bash tools/ncu_capture.sh TAG [FRAMES]
set -u
TAG=${1:-r2}; F=${2:-128}
OUT=gpurun_out
mkdir -p $OUT
python tools/prof_all.py $F > $OUT/${TAG}_profall.log 2>&1 || { tail -5 $OUT/${TAG}_profall.log; exit 1; }
# the first pass of prof_all.py is the warm-up: count its launches and skip them
N=$(ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prof_all.py $F 2>/dev/null | grep -c '"gpu__time_duration.sum"')
SKIP=$((N / 2))
echo "launches per run: $N, skipping $SKIP" > $OUT/${TAG}_ncu.log
ncu --set full --clock-control none --launch-skip $SKIP -o /tmp/${TAG}_all -f python tools/prof_all.py $F >> $OUT/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}_all.ncu-rep --page raw --csv > $OUT/${TAG}_all_raw.csv 2>/dev/null
ls -la /tmp/${TAG}_all.ncu-rep $OUT/${TAG}_all_raw.csv >> $OUT/${TAG}_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --pipeline 0 --no-cpu --no-e2e --no-latency --no-post > $OUT/${TAG}_ncu2.log 2>&1
tail -c 300 $OUT/${TAG}_ncu.log
