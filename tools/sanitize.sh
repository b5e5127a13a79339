#!/bin/bash
# compute-sanitizer passes over small parity runs of both device models (memcheck + racecheck + synccheck)
set -o pipefail
for tool in memcheck racecheck synccheck; do
  for t in tools/m2s_tiny.py tools/misp_tiny.py; do
    echo "== $tool $t"
    timeout 600 compute-sanitizer --tool $tool python $t 2>&1 | grep -E "ERROR SUMMARY|parity ok|RACECHECK SUMMARY|hazard|Error" | head -8
  done
done
