#!/bin/bash
# compute-sanitizer passes (memcheck + racecheck + synccheck) over small parity runs of both device models: the layer-by-layer kernels
# (incl. the two-class finish: DDO_FINISH_SPLIT_MIN=2 forces it on a tiny batch), the FRONTIER-cutset kernels, and the persistent
# whole-DD kernel (DDO_DD=1) at two cluster sizes.  usage: tools/sanitize.sh > profiles/rNN_sanitizer.txt
set -o pipefail
run() {  # run <label> <env...> -- <script>
  local label=$1; shift
  for tool in memcheck racecheck synccheck; do
    echo "== $tool | $label"
    env "${@:1:$#-1}" timeout 900 compute-sanitizer --tool $tool python "${@: -1}" 2>&1 | grep -E "ERROR SUMMARY|parity ok|RACECHECK SUMMARY|hazard|Error|error" | head -8
  done
}
run "MAX2SAT engine" X=1 tests/tools/m2s_tiny.py
run "MISP engine (lock-step kernels, FRONTIER kernels)" X=1 tests/tools/misp_tiny.py
run "MISP engine, two-class finish forced" DDO_FINISH_SPLIT_MIN=2 DDO_FINISH_CL_MAX=0 tests/tools/misp_tiny.py
run "MISP engine, persistent whole-DD kernel, single-CTA clusters" DDO_DD=1 DDO_DD_CS=1 tests/tools/misp_tiny.py
run "MISP engine, persistent whole-DD kernel, clusters of 4" DDO_DD=1 DDO_DD_CS=4 tests/tools/misp_tiny.py
