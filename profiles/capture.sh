#!/bin/bash
# usage (on the GPU box): profiles/capture.sh <kernel-regex> <tag> -- <command...>
# one `ncu --set full` capture of the first matching launch after 2 skipped ones; CSV pages land in gpurun_out/
KREGEX="$1"; TAG="$2"; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${KREGEX}" -s 2 -c 1 -f -o gpurun_out/${TAG} "$@" > gpurun_out/${TAG}.ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --print-source sass > gpurun_out/${TAG}_sass.csv 2>/dev/null
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_cuda.csv 2>/dev/null
rm -f gpurun_out/${TAG}.ncu-rep
tail -2 gpurun_out/${TAG}.ncu.log
