#!/bin/bash
# Full GPU validation: whole test-suite (all assets), bench lines per BASELINE configuration, reference arm.
mkdir -p gpurun_out
TAG=${TAG:-full}
echo "== all gpu tests"; timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/t_all_$TAG.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_all_$TAG.log
grep -E "max\|d\||golden file|ragged|invariance|stream kernel|bulk|batched server|TCP|edge input|reset|B=|cache|other|Hz" gpurun_out/t_all_$TAG.log | grep -v "frame \|max|d| vs reference fixture" > gpurun_out/parity_numbers_$TAG.txt; grep -E "golden file|bulk|FAILED|Error" gpurun_out/parity_numbers_$TAG.txt gpurun_out/t_all_$TAG.log | head -20
if [ -n "$BENCH" ]; then
echo "== bench config 2"; timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; tail -2 gpurun_out/bench_$TAG.err
for C in 3 4 5; do timeout 600 python bench.py --config $C --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_config$C.json 2>> gpurun_out/bench_$TAG.err; done
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err; echo "rc=$?"
python - <<PY
import json
for f in ["bench_$TAG.json", "bench_${TAG}_config3.json", "bench_${TAG}_config4.json", "bench_${TAG}_config5.json", "bench_${TAG}_reference.json"]:
    try:
        d = json.load(open("gpurun_out/" + f))
        print(f, "value", round(d["value"], 1), "ms/step", round(d.get("ms_per_step", 0), 4), "e2e", round(d["e2e"]["value"], 1), "launches/step", d.get("gpu_launches_per_step"),
              "roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.get("roofline", {}).items() if k in ("achieved", "frac", "frac_of_bf16x3_peak", "us_per_launch", "share_of_step")})
    except Exception as e:
        print(f, "ERR", e)
PY
fi
