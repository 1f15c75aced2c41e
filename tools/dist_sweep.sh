#!/bin/sh
# usage: dist_sweep.sh NGPUS  -- distributed correctness check, then the bench over chunk / comm-CTA settings
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
$TR --master-port 29533 tests/dist_gpu_check.py > gpurun_out/r2_dist_check_p$N.log 2>&1
echo "dist check rc=$? ok=$(grep -c ' OK' gpurun_out/r2_dist_check_p$N.log) fail=$(grep -c 'FAIL' gpurun_out/r2_dist_check_p$N.log)"
grep -v " OK" gpurun_out/r2_dist_check_p$N.log | tail -15
port=29600
for chunks in 1 4 8; do
  for ctas in 0 296 148 74; do
    if [ $chunks = 1 ] && [ $ctas != 0 ]; then continue; fi
    port=$((port+1))
    FFTW3_B200_DIST_CHUNKS=$chunks FFTW3_B200_DIST_COMM_CTAS=$ctas $TR --master-port $port bench.py --gpus $N --steps 5 --warmup 3 --estimate --no-e2e --no-cpu --no-check > gpurun_out/sweep_$N.json 2> gpurun_out/sweep_$N.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/sweep_$N.json").read().strip().splitlines()[-1])
    print("P=$N chunks=$chunks ctas=$ctas: %.3f ms natural, %.3f ms transposed-out, stages %s" % (d["ms_per_step"], d["config"]["transposed_out_ms_per_step"], d["roofline"]["nvlink"].get("stage_ms")))
except Exception as e:
    print("P=$N chunks=$chunks ctas=$ctas: FAILED", e); print(open("gpurun_out/sweep_$N.err").read()[-1500:])
PY
  done
done
