#!/bin/bash
# Round-end evidence on the GPU box: full GPU test-suite, bench lines (B = 64 headline, 256, 1024, reference arm),
# steady-state ncu launch list and one ncu --set full capture of the stream kernel.  TAG names the outputs.
mkdir -p gpurun_out
TAG=${TAG:-final}
echo "== tests"; timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/t_all_$TAG.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t_all_$TAG.log
grep -E "max\|d\||golden file|ragged|invariance" gpurun_out/t_all_$TAG.log | grep -v "frame " > gpurun_out/parity_numbers_$TAG.txt
echo "== bench B=64"; timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "rc=$?"; tail -2 gpurun_out/bench_$TAG.err
for B in 256 1024; do timeout 600 python bench.py --steps 50 --warmup 5 --batch-per-gpu $B --no-cpu-baseline > gpurun_out/bench_${TAG}_B$B.json 2>> gpurun_out/bench_$TAG.err; done
timeout 600 python bench.py --steps 50 --warmup 5 --batch-per-gpu 64 --ctx-frames 100 --no-cpu-baseline > gpurun_out/bench_${TAG}_T100.json 2>> gpurun_out/bench_$TAG.err
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err; echo "rc=$?"
python - <<PY
import json
for f in ["bench_$TAG.json", "bench_${TAG}_B256.json", "bench_${TAG}_B1024.json", "bench_${TAG}_T100.json", "bench_${TAG}_reference.json"]:
    try:
        d = json.load(open("gpurun_out/" + f))
        print(f, "value", round(d["value"], 1), "ms/step", round(d.get("ms_per_step", 0), 4), "e2e", round(d["e2e"]["value"], 1), "launches/step", d.get("gpu_launches_per_step"),
              "roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.get("roofline", {}).items() if k in ("achieved", "frac", "frac_of_bf16x3_peak", "us_per_launch", "share_of_step")})
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== ncu launch list (steady state: window full)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1270 -c 92 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "rc=$?"
python tools/agg_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_agg.txt; head -12 gpurun_out/launches_${TAG}_agg.txt
echo "== ncu full capture of the stream kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_tf -s 60 -c 1 -o gpurun_out/prof_stream_tf_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "rc=$?"
