#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-alias}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -q -s -k "stream_kernel or golden or window or ragged or at_size or cache" > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_$TAG.log; grep -E "golden file \(conv4p=1 \{\}\)" gpurun_out/t_$TAG.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2>/dev/null; python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('step', round(d['ms_per_step'],4), round(d['value']), 'stream us', d['roofline']['us_per_launch'])"
DBG_OP=12 FUSED_V=2 timeout 300 python tools/fused_clocks.py > gpurun_out/fused_clocks_$TAG.log 2>&1; grep ^total gpurun_out/fused_clocks_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_stream_tf2 -s 60 -c 3 --csv --log-file gpurun_out/ncu_dram_$TAG.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dram_$TAG.log 2>&1; echo "ncu rc=$?"; grep -E "k_stream_tf2" gpurun_out/ncu_dram_$TAG.csv | awk -F'","' '{print $(NF-3), $(NF-2), $(NF-1), $NF}' | head -12
