#!/bin/bash
# A/B of whole-step time at a BASELINE configuration for library variants on ONE box, interleaved.
mkdir -p gpurun_out
TAG=${TAG:-abc}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "stream_kernel" > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_$TAG.log
for R in $(seq 1 ${REPS:-2}); do
  for V in base ${VARIANTS}; do
    L=$PWD/vap_realtime_b200/libvapb200_$V.so; [ $V = base ] && L=$PWD/vap_realtime_b200/libvapb200.so
    for C in ${CONFIGS:-5}; do
    VAPB_LIB=$L timeout 300 python bench.py --config $C --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_${V}_c${C}_$R.json 2>/dev/null
    echo "$V config $C rep $R: step $(python -c "import json;d=json.load(open('gpurun_out/bench_${TAG}_${V}_c${C}_$R.json'));print(round(d['ms_per_step'],4), round(d['value']), 'stream us', d['roofline'].get('us_per_launch'))")"
    done
  done
done
