#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-tail}
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -q -s -x -k "${KEXPR:-not selftest and not golden_full}" > gpurun_out/t_$TAG.log 2>&1; echo "rc=$?"
grep -E "max\|d\||stream kernel|ragged|passed|failed|FAILED|Error|error|timed out|B=" gpurun_out/t_$TAG.log | grep -v "frame " | tail -45
echo "== ablation"; ONLY=${ONLY:-default,no_tail,stream_v2,lstm_x_tc,no_fused} timeout 600 python tools/step_ablation.py > gpurun_out/ablation_$TAG.log 2>&1; echo "rc=$?"; cat gpurun_out/ablation_$TAG.log
