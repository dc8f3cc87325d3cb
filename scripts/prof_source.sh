#!/bin/bash
# ncu capture of the dominant kernel with per-instruction (SASS) sampling, exported as gzip'd CSV:
#   gpurun --timeout 900 -- 'bash scripts/prof_source.sh [TC_VERSION]'
# Output: gpurun_out/src_<tag>.csv.gz (source page, SASS view), gpurun_out/src_<tag>_raw.csv (raw page).
set -u
cd "$(dirname "$0")/.."
VER=${1:-3}
TAG="v${VER}"
mkdir -p gpurun_out
export B200KGE_TC_VERSION=$VER B200KGE_TC4_DIRECT=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pairwise_tc -s 4 -c 1 \
  -o /tmp/psrc_$TAG python bench.py --steps 4 --warmup 3 > gpurun_out/src_$TAG.log 2>&1
timeout 120 ncu -i /tmp/psrc_$TAG.ncu-rep --page raw --csv > gpurun_out/src_${TAG}_raw.csv 2>> gpurun_out/src_$TAG.log
timeout 200 ncu -i /tmp/psrc_$TAG.ncu-rep --page source --csv --print-source sass 2>> gpurun_out/src_$TAG.log | gzip > gpurun_out/src_${TAG}.csv.gz
ls -la gpurun_out | grep src_
