#!/bin/bash
# Quick GPU iteration: targeted parity tests, per-op clocks of the stream kernel, option ablation.  TAG names the outputs.
mkdir -p gpurun_out
TAG=${TAG:-q}
echo "== parity"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "${KEXPR:-not selftest and not golden_file}" > gpurun_out/t_quick_$TAG.log 2>&1; echo "rc=$?"
grep -E "max\|d\||golden file|ragged|invariance|passed|failed|FAILED|Error|error|timed out" gpurun_out/t_quick_$TAG.log | grep -v "frame " | tail -40
echo "== fused clocks"
timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_$TAG.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/fused_clocks_$TAG.log
echo "== ablation"
ONLY=${ONLY:-default,no_fused} timeout 600 python tools/step_ablation.py > gpurun_out/ablation_$TAG.log 2>&1; echo "rc=$?"; cat gpurun_out/ablation_$TAG.log
