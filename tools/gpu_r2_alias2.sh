#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-alias2}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py tests/test_gpu_dropin.py -m gpu -q -s > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_$TAG.log; grep -E "golden file \(conv4p=1 \{|T=100|T=128|B=256" gpurun_out/t_$TAG.log | head
for C in 2 3 4 5; do timeout 300 python bench.py --config $C --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_c$C.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/bench_${TAG}_c$C.json'));print('config $C step', round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']), 'stream us', d['roofline'].get('us_per_launch'))"; done
