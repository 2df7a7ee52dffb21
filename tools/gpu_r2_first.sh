#!/bin/bash
# Round 2, first GPU contact: the new at-size tests, the whole GPU suite, a bench line, sanitizer runs.
mkdir -p gpurun_out
TAG=${TAG:-r2a}
echo "== at-size tests"; timeout 900 python -m pytest tests/test_gpu_at_size.py -m gpu -q -s -x > gpurun_out/t_size_$TAG.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_size_$TAG.log
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/t_all_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_all_$TAG.log
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; tail -2 gpurun_out/bench_$TAG.err
python - <<PY
import json
for f in ["bench_$TAG.json"]:
    try:
        d = json.load(open("gpurun_out/" + f))
        print(f, "value", round(d["value"], 1), "ms/step", round(d.get("ms_per_step", 0), 4), "e2e", round(d["e2e"]["value"], 1), "launches/step", d.get("gpu_launches_per_step"))
        print({k: v for k, v in d["kernel_breakdown_ms"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== sanitizer"
for F in 2 0; do
  for TOOL in memcheck racecheck; do
    timeout 600 compute-sanitizer --tool $TOOL python tools/sanitize.py --fused $F --steps 4 > gpurun_out/sanitizer_${TOOL}_fused${F}_$TAG.log 2>&1; echo "$TOOL fused=$F rc=$?"; tail -4 gpurun_out/sanitizer_${TOOL}_fused${F}_$TAG.log
  done
done
