#!/bin/bash
# A/B: programmatic dependent launch inside the step graph (library built with -DVAPB_ENABLE_PDL, option pdl=1) vs plain edges
mkdir -p gpurun_out
TAG=${TAG:-pdl}
for R in 1 2; do
  for V in base pdl1 pdl0; do
    L=$PWD/vap_realtime_b200/libvapb200_pdl.so; O="--opt pdl=1"; [ $V = base ] && L=$PWD/vap_realtime_b200/libvapb200.so && O=""; [ $V = pdl0 ] && O="--opt pdl=0"
    VAPB_LIB=$L timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline $O > gpurun_out/bench_${TAG}_${V}_$R.json 2>gpurun_out/bench_${TAG}_${V}_$R.err
    echo "$V rep $R: $(python -c "import json;d=json.load(open('gpurun_out/bench_${TAG}_${V}_$R.json'));print(round(d['ms_per_step'],4), round(d['value']), 'b2b', round(d['back_to_back_ms_per_step'],4))")"
  done
done
