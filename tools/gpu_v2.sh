#!/bin/bash
# Stream kernel v2 bring-up: edge tests, fixture parity, per-op clocks (v2 and v1), bench.
mkdir -p gpurun_out
TAG=${TAG:-v2a}
echo "== window edges + fixture"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -x -k "${KEXPR:-window_edges or fixture_ctx2500 or ragged or option_variants}" > gpurun_out/t_v2_$TAG.log 2>&1; echo "rc=$?"
grep -E "max\|d\||stream kernel|ragged|passed|failed|FAILED|Error|error|timed out" gpurun_out/t_v2_$TAG.log | tail -40
echo "== clocks v2"; FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}.log 2>&1; echo "rc=$?"; tail -32 gpurun_out/fused_clocks_${TAG}.log
if [ -n "$FULL" ]; then
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/t_all_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_all_$TAG.log
fi
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 5 ${BENCH_ARGS:---no-cpu-baseline} > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$TAG.json"))
    print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "launches/step", d.get("gpu_launches_per_step"))
    print({k: v["ms"] for k, v in d["kernel_breakdown_ms"].items()})
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.get("roofline", {}).items() if k in ("achieved", "frac", "frac_of_bf16x3_peak", "us_per_launch", "share_of_step")})
except Exception as e:
    print("ERR", e)
PY
