#!/bin/bash
# A/B of stream-kernel builds on ONE box: every variant REPS times, interleaved; prints the totals.
mkdir -p gpurun_out
TAG=${TAG:-ab}
for R in $(seq 1 ${REPS:-3}); do
  for V in base ${VARIANTS}; do
    L=$PWD/vap_realtime_b200/libvapb200_$V.so; [ $V = base ] && L=$PWD/vap_realtime_b200/libvapb200.so
    VAPB_LIB=$L DBG_OP=${DBG_OP:-12} FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}_${V}_$R.log 2>&1
    echo "$V rep $R: $(grep ^total gpurun_out/fused_clocks_${TAG}_${V}_$R.log)"
  done
done
