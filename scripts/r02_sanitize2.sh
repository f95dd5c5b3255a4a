#!/bin/bash
# compute-sanitizer over the kernels of the second half of round 2 (K1s, split reductions, staged NUTS)
mkdir -p gpurun_out
export SANITIZE_SITES=97 SANITIZE_SITES_BIG=40000
for tool in memcheck synccheck; do
  echo "== $tool: small"
  timeout 500 compute-sanitizer --tool $tool python scripts/sanitize.py small 2>&1 | grep -E "ok$|plan|ERROR SUMMARY|Error|error" | tail -8
done
export SANITIZE_SITES=65 SANITIZE_SITES_BIG=20000
echo "== racecheck: small"
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize.py small 2>&1 | grep -E "ok$|plan|RACECHECK SUMMARY|hazard|Error" | tail -8
