#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-v2h}
echo "== parity quick"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "window_edges or fixture_ctx2500 or ragged or option_variants" > gpurun_out/t_v2_$TAG.log 2>&1; echo "rc=$?"
grep -E "max\|d\||stream kernel|passed|failed|FAILED|Error|error|timed out" gpurun_out/t_v2_$TAG.log | tail -30
for V in 2 1; do
echo "== clocks v$V"; DBG_OP=12 FUSED_V=$V timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}_v$V.log 2>&1; echo "rc=$?"; grep -E "ffn1|total" gpurun_out/fused_clocks_${TAG}_v$V.log
done
