#!/bin/bash
# compute-sanitizer memcheck + synccheck on the final stream kernels (racecheck of these kernels does not terminate, see profiles/r02_n_*)
mkdir -p gpurun_out
TAG=${TAG:-san2}
run() {  # tool, name, args...
  local tool=$1 name=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize.py "$@" > gpurun_out/sanitizer_${tool}_${name}_$TAG.log 2>&1
  echo "$tool $name rc=$?"; grep -E "ERROR SUMMARY|sanitize target|Error:|hazard" gpurun_out/sanitizer_${tool}_${name}_$TAG.log | sort | uniq -c | sort -rn | head -6
}
run memcheck stream_v2 --fused 2 --fused_v 2 --steps 4 --B 3
run memcheck stream_v1 --fused 2 --fused_v 1 --steps 4 --B 3
run synccheck stream_v2 --fused 2 --fused_v 2 --steps 3 --B 3
