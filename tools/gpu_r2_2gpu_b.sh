#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-2gpub}
echo "== nccl test"; timeout 600 python -m pytest tests/test_gpu_dist_nccl.py -m gpu -q -s > gpurun_out/t_nccl_$TAG.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_nccl_$TAG.log
echo "== bench 2 GPUs config 2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --config 2 --steps 200 --warmup 5 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "rc=$?"; tail -3 gpurun_out/bench_${TAG}_n2.err
echo "== bench 1 GPU config 2 (same box)"; timeout 600 python bench.py --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "rc=$?"
python - <<PY
import json
for f in ["bench_${TAG}_n2.json", "bench_${TAG}_n1.json"]:
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().split("\n")[-1])
        print(f, "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d.get("ms_per_step", 0), 4), "e2e", round(d["e2e"]["value"], 1), "b2b", d.get("back_to_back_ms_per_step"), d["config"]["global_streams"])
    except Exception as e:
        print(f, "ERR", e)
PY
