#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-var}
for V in ${VARIANTS}; do
  echo "== variant $V (timing only)"; VAPB_LIB=$PWD/vap_realtime_b200/libvapb200_$V.so DBG_OP=${DBG_OP:-12} FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}_$V.log 2>&1; echo "rc=$?"; tail -${TAILN:-34} gpurun_out/fused_clocks_${TAG}_$V.log
done
