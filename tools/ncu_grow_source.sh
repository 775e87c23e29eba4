#!/bin/bash
# Runs ON THE GPU BOX: source-level ncu capture (stall samples per SASS / CUDA line) of the region-growing kernel for F frames.
# Keeps the report in /tmp, brings back the source page as CSV.  Usage: bash tools/ncu_grow_source.sh TAG F
set -u
TAG=$1; F=$2
mkdir -p gpurun_out
python tools/prof_one.py $F > gpurun_out/${TAG}_run.log 2>&1 || { tail -5 gpurun_out/${TAG}_run.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:k_lsd_grow2 --launch-skip 1 -c 1 -o /tmp/${TAG} -f python tools/prof_one.py $F > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i /tmp/${TAG}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_source.csv 2>/dev/null
ncu -i /tmp/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_source.csv /tmp/${TAG}.ncu-rep
tail -3 gpurun_out/${TAG}_ncu.log
