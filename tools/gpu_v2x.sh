#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-v2x}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "stream or golden or window" > gpurun_out/t_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_$TAG.log; grep -E "golden file|stream kernel" gpurun_out/t_$TAG.log | head
echo "== base"; DBG_OP=${DBG_OP:-12} FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}_base.log 2>&1; tail -34 gpurun_out/fused_clocks_${TAG}_base.log
TAILN=${TAILN:-11} VARIANTS="$VARIANTS" TAG=$TAG bash tools/gpu_v2var.sh
for V in ${TESTVARIANTS}; do
  echo "== parity of variant $V"; VAPB_LIB=$PWD/vap_realtime_b200/libvapb200_$V.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "stream_kernel or golden" > gpurun_out/t_${TAG}_$V.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t_${TAG}_$V.log; grep -E "golden file|stream kernel v2 T=50" gpurun_out/t_${TAG}_$V.log | head
done
