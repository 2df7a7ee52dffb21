#!/bin/bash
# compute-sanitizer on the hot path: memcheck with the product library, racecheck with the unbounded-wait build
# (the instrumented producer warps are orders of magnitude slower than the polling warps).
mkdir -p gpurun_out
TAG=${TAG:-san}
run() {  # tool, lib, name, args...
  local tool=$1 lib=$2 name=$3; shift 3
  VAPB_LIB=$lib timeout 900 compute-sanitizer --tool $tool python tools/sanitize.py "$@" > gpurun_out/sanitizer_${tool}_${name}_$TAG.log 2>&1
  echo "$tool $name rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target|Error:|hazard" gpurun_out/sanitizer_${tool}_${name}_$TAG.log | sort | uniq -c | sort -rn | head -6
}
L=$PWD/vap_realtime_b200/libvapb200.so
S=$PWD/vap_realtime_b200/libvapb200_san.so
run memcheck $L stream_v1 --fused 2 --fused_v 1 --steps 4 --B 3
run memcheck $L stream_v2 --fused 2 --fused_v 2 --steps 4 --B 3
run memcheck $L batched_bulk --fused 0 --steps 4 --B 3 --bulk 12
run racecheck $S stream_v1 --fused 2 --fused_v 1 --steps 2 --B 1
run racecheck $S stream_v2 --fused 2 --fused_v 2 --steps 2 --B 1
run racecheck $S batched --fused 0 --steps 2 --B 1
