#!/bin/bash
# `ncu --set full` on the launches of ONE steady-state step whose kernel name matches $1 (regex), at most $2 launches;
# writes the details page (text) and the raw page (csv) to gpurun_out/ncu_<tag>.{txt,csv}; the report stays in /tmp.
PATTERN="$1"; COUNT="${2:-20}"; TAG="${3:-sel}"
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profile-range --no-graph"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$PATTERN" -c "$COUNT" \
    -f -o /tmp/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "capture exit=$?"
ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/ncu_$TAG.txt 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_$TAG.csv 2>/dev/null
ls -la gpurun_out/ncu_$TAG.*
