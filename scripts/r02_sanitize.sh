#!/bin/bash
mkdir -p gpurun_out
export SANITIZE_SITES=97
for tool in memcheck synccheck; do
  echo "== $tool (97 sites): occu occu_rn extras" 
  timeout 500 compute-sanitizer --tool $tool python scripts/sanitize.py occu occu_rn extras 2>&1 | grep -E "ok$|ERROR SUMMARY|Error|error" | tail -6
done
export SANITIZE_SITES=65
echo "== racecheck (65 sites): occu occu_rn"
timeout 600 compute-sanitizer --tool racecheck python scripts/sanitize.py occu occu_rn 2>&1 | grep -E "ok$|RACECHECK SUMMARY|hazard|Error" | tail -6
