#!/bin/bash
# Iteration script for the GPU box: parity (product path), bench, ncu launch list.  TAG names the outputs.
mkdir -p gpurun_out
TAG=${TAG:-iter}
echo "== selftest"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k selftest > gpurun_out/t_selftest_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_selftest_$TAG.log
echo "== parity (both engines)"
timeout 1800 python -m pytest tests -m gpu -q -s -k "not selftest" > gpurun_out/t_parity_$TAG.log 2>&1; echo "rc=$?"; grep -E "max\|d\||golden file|ragged|invariance|passed|failed|FAILED|Error" gpurun_out/t_parity_$TAG.log | grep -v "frame " | tail -40
echo "== bench"
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value", d["value"], "ms/step", d["ms_per_step"], "b2b", d["back_to_back_ms_per_step"], "e2e", d["e2e"]["value"], "frac_x3", d["roofline"]["frac_of_bf16x3_peak"], "launches", d["gpu_launches_per_step"])
print({k:v["ms"] for k,v in d["kernel_breakdown_ms"].items()})
print(d.get("cpu_baseline"))
PY
for B in 256 1024; do
  timeout 600 python bench.py --steps 50 --warmup 5 --batch-per-gpu $B --no-cpu-baseline > gpurun_out/bench_${TAG}_B$B.json 2>> gpurun_out/bench_$TAG.err
  python -c "import json; d=json.load(open('gpurun_out/bench_${TAG}_B$B.json')); print('B=$B value', d['value'], 'ms/step', d['ms_per_step'], 'frac_x3', d['roofline']['frac_of_bf16x3_peak'])"
done
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "rc=$?"
