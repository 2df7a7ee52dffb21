#!/bin/bash
# 1 -> N GPU weak-scaling lines on ONE box (driver contract launch), configs 2 and 3 at the largest N.
mkdir -p gpurun_out
TAG=${TAG:-scale}
NMAX=${NMAX:-8}
P=29530
for N in $NMAX 4 2; do
  [ $N -gt $NMAX ] && continue
  for C in 2 $( [ $N = $NMAX ] && echo 3 ); do
    P=$((P+1))
    echo "== bench $N GPUs config $C"; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config $C --steps 100 --warmup 5 > gpurun_out/bench_${TAG}_n${N}_config$C.json 2> gpurun_out/bench_${TAG}_n${N}_config$C.err; echo "rc=$?"; tail -2 gpurun_out/bench_${TAG}_n${N}_config$C.err
  done
done
echo "== bench 1 GPU config 2 (same box)"; timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1_config2.json 2> gpurun_out/bench_${TAG}_n1_config2.err; echo "rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_n*_config*.json")):
    try:
        d = json.loads(open(f).read().strip().split("\n")[-1])
        print(f, "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d.get("ms_per_step", 0), 4), "e2e", round(d["e2e"]["value"], 1), "b2b", d.get("back_to_back_ms_per_step"), d["config"].get("global_streams"))
    except Exception as e:
        print(f, "ERR", e)
PY
