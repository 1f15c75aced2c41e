#!/bin/sh
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
port=29700
for cfg in "8 296 0" "8 296 1" "8 444 0" "4 296 0" "16 296 0" "8 592 0"; do
  set -- $cfg; chunks=$1; ctas=$2; full=$3
  port=$((port+1))
  if [ "$full" = "1" ]; then export FFTW3_B200_DIST_X_FULLREGS=1; else unset FFTW3_B200_DIST_X_FULLREGS; fi
  FFTW3_B200_DIST_CHUNKS=$chunks FFTW3_B200_DIST_COMM_CTAS=$ctas $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/sweep2.json 2> gpurun_out/sweep2.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep2.json").read().strip().splitlines()[-1])
    print("chunks=$chunks ctas=$ctas fullregs=$full: %.3f ms natural, %.3f ms transposed-out, stages %s" % (d["ms_per_step"], d["config"]["transposed_out_ms_per_step"], d["roofline"]["nvlink"].get("stage_ms")))
except Exception as e:
    print("chunks=$chunks ctas=$ctas: FAILED", e); print(open("gpurun_out/sweep2.err").read()[-800:])
PY
done
