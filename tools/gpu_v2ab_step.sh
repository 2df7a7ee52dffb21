#!/bin/bash
# A/B of whole-step time (bench.py, config 2) and in-kernel clocks for library variants on ONE box, interleaved.
mkdir -p gpurun_out
TAG=${TAG:-abs}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "stream_kernel or golden or window or ragged" > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_$TAG.log
for R in $(seq 1 ${REPS:-2}); do
  for V in base ${VARIANTS}; do
    L=$PWD/vap_realtime_b200/libvapb200_$V.so; [ $V = base ] && L=$PWD/vap_realtime_b200/libvapb200.so
    VAPB_LIB=$L DBG_OP=0 FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_${TAG}_${V}_$R.log 2>&1
    VAPB_LIB=$L timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_${V}_$R.json 2>/dev/null
    echo "$V rep $R: $(grep ^gather_ring gpurun_out/fused_clocks_${TAG}_${V}_$R.log) $(grep ^total gpurun_out/fused_clocks_${TAG}_${V}_$R.log) | step $(python -c "import json;d=json.load(open('gpurun_out/bench_${TAG}_${V}_$R.json'));print(round(d['ms_per_step'],4), round(d['value']))")"
  done
done
