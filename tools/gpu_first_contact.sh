#!/bin/bash
# First-contact script for the GPU box: bisects failures by GEMM engine so one broken kernel
# cannot hide the state of the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/env.log 2>&1
ls /root/reference >> gpurun_out/env.log 2>&1
nproc >> gpurun_out/env.log; lscpu | grep "Model name" >> gpurun_out/env.log
echo "== selftest (tcgen05 gemm)" 
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k selftest > gpurun_out/t_selftest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_selftest.log
echo "== parity, fp32 CUDA-core GEMMs"
VAPB_TEST_GEMM=0 timeout 1500 python -m pytest tests -m gpu -q -s -k "not selftest" > gpurun_out/t_gemm0.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_gemm0.log
echo "== parity, tcgen05 GEMMs"
VAPB_TEST_GEMM=1 timeout 1500 python -m pytest tests -m gpu -q -s -k "not selftest" > gpurun_out/t_gemm1.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_gemm1.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench gemm=0"
timeout 600 python bench.py --steps 50 --warmup 5 --gemm 0 --no-cpu-baseline > gpurun_out/bench_gemm0.json 2> gpurun_out/bench_gemm0.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_gemm0.json
echo "== bench gemm=1"
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_gemm1.json 2> gpurun_out/bench_gemm1.err; echo "rc=$?"; tail -c 3000 gpurun_out/bench_gemm1.json; tail -5 gpurun_out/bench_gemm1.err
