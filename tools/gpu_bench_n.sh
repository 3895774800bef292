#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it; result to gpurun_out/r03/bench_n$N.json. usage: bash tools/gpu_bench_n.sh N
N=$1; O=gpurun_out/r03; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo rc=$?
tail -2 $O/bench_n$N.err | cut -c1-200
python - $N <<'PY'
import json, sys
n = sys.argv[1]
l = json.loads(open(f"gpurun_out/r03/bench_n{n}.json").read().strip().splitlines()[-1])
oc = l.get("one_capture") or {}
print("value", l["value"], "sustained", (l.get("sustained") or {}).get("value"), "e2e", l["e2e"]["value"], "ceiling", l["e2e"]["copy_ceiling"]["value"])
print("one_capture", oc.get("value"), oc.get("mode"), {k: (v or {}).get("value") for k, v in oc.items() if isinstance(v, dict)})
PY
