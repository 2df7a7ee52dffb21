#!/bin/bash
# Validation of the wave-fill path rule: at-size tests, the cache tests, bench configs 3 / 4 / 5.
mkdir -p gpurun_out
TAG=${TAG:-paths}
timeout 900 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_parity.py -m gpu -q -s -k "at_size or cache or big or config or ragged or window" > gpurun_out/t_paths_$TAG.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_paths_$TAG.log
grep -E "max\|d\||B=" gpurun_out/t_paths_$TAG.log | head -30
for C in 3 4 5; do timeout 600 python bench.py --config $C --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_config$C.json 2>> gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}_config$C.json')); print($C, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1), d.get('gpu_launches_per_step'))"; done
