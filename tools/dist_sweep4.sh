#!/bin/sh
# usage: dist_sweep4.sh NGPUS "split chunks" ...  -- copies per block (streams / copy engines) and chunk counts
N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29900
for cfg in "$@"; do
  set -- $cfg; split=$1; chunks=$2
  port=$((port+1))
  FFTW3_B200_DIST_COPY_SPLIT=$split FFTW3_B200_DIST_CHUNKS=$chunks $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/sweep4.json 2> gpurun_out/sweep4.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep4.json").read().strip().splitlines()[-1])
    print("split=$split chunks=$chunks: %.3f ms natural, %.3f ms transposed-out, stages %s" % (d["ms_per_step"], d["config"]["transposed_out_ms_per_step"], d["roofline"]["nvlink"].get("stage_ms")))
except Exception as e:
    print("split=$split chunks=$chunks: FAILED", e); print(open("gpurun_out/sweep4.err").read()[-1500:])
PY
done
