#!/bin/bash
# ncu capture of the dominant kernel and the launch list for one tensor-core path (B200KGE_TC3_TK from the environment):
#   gpurun --timeout 900 -- 'bash scripts/exp_profile.sh 3'        # B200KGE_TC_VERSION (1 default, 3, 4)
#   gpurun --timeout 900 -- 'bash scripts/exp_profile.sh 4 1'      # tc4 with direct TMA signalling
# Reports stay in /tmp on the box (they exceed gpurun_out's size limit); the raw-page CSV and the launch list
# come back in gpurun_out/ as prof_v<ver>[_d<direct>]_{raw,launches}.csv.  Numbers printed under ncu are not
# bench values.
set -u
cd "$(dirname "$0")/.."
VER=${1:-1}; DIRECT=${2:-0}
TAG="v${VER}_d${DIRECT}"
mkdir -p gpurun_out
export B200KGE_TC_VERSION=$VER B200KGE_TC4_DIRECT=$DIRECT
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pairwise_tc -s 4 -c 1 \
  -o /tmp/prof_$TAG python bench.py --steps 4 --warmup 3 > gpurun_out/prof_$TAG.log 2>&1
timeout 120 ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>> gpurun_out/prof_$TAG.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv \
  --log-file gpurun_out/prof_${TAG}_launches.csv python bench.py --steps 8 --warmup 3 >> gpurun_out/prof_$TAG.log 2>&1
timeout 200 python bench.py > gpurun_out/prof_${TAG}_bench.json 2>> gpurun_out/prof_$TAG.log
tail -3 gpurun_out/prof_$TAG.log; ls -la gpurun_out | tail -8
