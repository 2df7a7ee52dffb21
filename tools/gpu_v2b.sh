#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-v2b}
echo "== parity quick"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "window_edges or fixture_ctx2500 or ragged" > gpurun_out/t_v2_$TAG.log 2>&1; echo "rc=$?"
grep -E "max\|d\||stream kernel|passed|failed|FAILED|Error|error|timed out" gpurun_out/t_v2_$TAG.log | tail -14
for OP in ${OPS:-6 8 12 13}; do
  echo "== clocks v2 op $OP"; DBG_OP=$OP FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}_op$OP.log 2>&1; echo "rc=$?"
  if [ $OP = 6 ]; then head -25 gpurun_out/fused_clocks_${TAG}_op$OP.log; fi
  tail -10 gpurun_out/fused_clocks_${TAG}_op$OP.log
done
for V in ${VARIANTS:-nostore}; do
  echo "== variant $V (timing only)"; VAPB_LIB=$PWD/vap_realtime_b200/libvapb200_$V.so DBG_OP=6 FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}_$V.log 2>&1; echo "rc=$?"; tail -36 gpurun_out/fused_clocks_${TAG}_$V.log
done
