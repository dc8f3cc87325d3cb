#!/bin/bash
# Validate and time the EXPERIMENTAL tensor-core paths (B200KGE_TC_VERSION=3|4) on a B200:
#   gpurun --timeout 900 -- 'bash scripts/exp_check.sh'
# Every step runs under its own `timeout` so that a hanging kernel ends with its process, not with the box.
# Results land in gpurun_out/exp_*.log / exp_summary.txt / bench_v*.json.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/exp_summary.txt
export B200KGE_EXPERIMENTAL=1
timeout 180 python -m pytest tests/test_gpu_experimental.py -q -k "evaluator or job or reciprocal or penalties or csr" > gpurun_out/exp_evaluator.log 2>&1
echo "evaluator + job-trace pytest rc=$?" >> gpurun_out/exp_summary.txt
tail -4 gpurun_out/exp_evaluator.log >> gpurun_out/exp_summary.txt
timeout 180 python -m pytest tests/test_gpu_experimental.py -x -q -k "x_gemm" > gpurun_out/exp_gemm.log 2>&1
echo "x_gemm pytest rc=$?" >> gpurun_out/exp_summary.txt
timeout 240 python -m pytest tests/test_gpu_experimental.py -q -k "x_backward or x_ns_backward" > gpurun_out/exp_backward.log 2>&1
echo "x_backward pytest rc=$?" >> gpurun_out/exp_summary.txt
tail -5 gpurun_out/exp_backward.log >> gpurun_out/exp_summary.txt
for v in "tc3 and not tk32" tc3-tk32 tc4-forward tc4-direct; do
  f="gpurun_out/exp_$(echo "$v" | tr ' ' '_').log"
  timeout 180 python -m pytest tests/test_gpu_experimental.py -x -q -k "$v" > "$f" 2>&1
  echo "$v pytest rc=$?" >> gpurun_out/exp_summary.txt
  tail -3 "$f" >> gpurun_out/exp_summary.txt
done
for cfg in "1 0 64" "3 0 64" "3 0 32" "4 0 64" "4 1 64"; do
  set -- $cfg
  B200KGE_TC_VERSION=$1 B200KGE_TC4_DIRECT=$2 B200KGE_TC3_TK=$3 timeout 180 python bench.py --steps 100 --warmup 5 \
    > "gpurun_out/bench_v$1_d$2_k$3.json" 2> "gpurun_out/bench_v$1_d$2_k$3.err"
  echo "bench v$1 direct=$2 tk=$3 rc=$?" >> gpurun_out/exp_summary.txt
  python - "gpurun_out/bench_v$1_d$2_k$3.json" >> gpurun_out/exp_summary.txt <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  ms_per_step", j["ms_per_step"], "value", j["value"], "kernel_ms", j["roofline"].get("kernel_ms"), "e2e", j["e2e"]["value"])
except Exception as ex:
    print("  no bench line:", ex)
PY
done
cat gpurun_out/exp_summary.txt
