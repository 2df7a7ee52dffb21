#!/bin/bash
# ncu evidence of the final round-2 step: steady-state launch list, --set full of the stream kernel and of one whole step.
mkdir -p gpurun_out
TAG=${TAG:-s}
echo "== ncu launch list (steady state: window full)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 56 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "rc=$?"
python tools/agg_launches.py gpurun_out/launches_$TAG.csv > gpurun_out/launches_${TAG}_agg.txt; head -20 gpurun_out/launches_${TAG}_agg.txt
echo "== ncu full capture of the stream kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_tf2 -s 60 -c 1 -o gpurun_out/prof_stream_tf2_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "rc=$?"
echo "== ncu full capture of one whole steady-state step"
timeout 900 ncu --set full --clock-control none -s 812 -c 14 -o gpurun_out/prof_step_$TAG python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_step_$TAG.log 2>&1; echo "rc=$?"
ls -la gpurun_out | grep prof_
