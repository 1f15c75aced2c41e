#!/bin/sh
# usage: dist_sweep5.sh NGPUS "ENV=val ENV=val" ...  -- one bench run per environment string
N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=30000
for cfg in "$@"; do
  port=$((port+1))
  env $cfg $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/sweep5.json 2> gpurun_out/sweep5.err
  CFG="$cfg" python - <<'PY'
import json, os
cfg = os.environ["CFG"]
try:
    d = json.loads(open("gpurun_out/sweep5.json").read().strip().splitlines()[-1])
    print("%s: %.3f ms natural, %.3f ms transposed-out, stages %s" % (cfg, d["ms_per_step"], d["config"]["transposed_out_ms_per_step"], d["roofline"]["nvlink"].get("stage_ms")))
except Exception as e:
    print(cfg, ": FAILED", e); print(open("gpurun_out/sweep5.err").read()[-1500:])
PY
done
