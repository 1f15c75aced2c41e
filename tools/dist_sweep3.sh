#!/bin/sh
# usage: dist_sweep3.sh NGPUS  -- first exchange by copy engines vs fused stores, and chunk counts
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29800
for cfg in "copy 8" "stores 8" "copy 4" "copy 16" "copy 32"; do
  set -- $cfg; mode=$1; chunks=$2
  port=$((port+1))
  FFTW3_B200_DIST_EXCHANGE=$mode FFTW3_B200_DIST_CHUNKS=$chunks $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/sweep3.json 2> gpurun_out/sweep3.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep3.json").read().strip().splitlines()[-1])
    print("exchange=$mode chunks=$chunks: %.3f ms natural, %.3f ms transposed-out, stages %s, check %s" % (d["ms_per_step"], d["config"]["transposed_out_ms_per_step"], d["roofline"]["nvlink"].get("stage_ms"), (d.get("check") or {}).get("ok")))
except Exception as e:
    print("exchange=$mode chunks=$chunks: FAILED", e); print(open("gpurun_out/sweep3.err").read()[-1500:])
PY
done
