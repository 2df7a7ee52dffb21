#!/bin/bash
# Bench + ncu evidence on the GPU box.  Logs land in gpurun_out/.
mkdir -p gpurun_out
R=${ROUND:-r01}
echo "== graph-equals-eager test"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "graph_equals" 2>&1 | tail -3
echo "== bench gemm=0 (fp32 CUDA-core GEMMs)"
timeout 900 python bench.py --steps 50 --warmup 5 --gemm 0 --no-cpu-baseline > gpurun_out/bench_${R}_gemm0.json 2> gpurun_out/bench_gemm0.err; echo "rc=$?"; tail -3 gpurun_out/bench_gemm0.err
echo "== bench gemm=1 (product path)"
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/bench_${R}_gemm1.json 2> gpurun_out/bench_gemm1.err; echo "rc=$?"; tail -3 gpurun_out/bench_gemm1.err
cat gpurun_out/bench_${R}_gemm1.json
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_${R}_reference.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_${R}_reference.json
echo "== ncu launch list (same bench command, 2 steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc=$?"
echo "== ncu full capture of the tcgen05 GEMM (conv1 + ffn shapes)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 40 -c 6 -o gpurun_out/prof_gemm_tc_${R} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
