#!/bin/sh
# usage: dist_final.sh NGPUS [check]  -- (optional) distributed correctness check, then both bench arms at N GPUs
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
if [ "$2" = "check" ]; then
  $TR --master-port 29533 tests/dist_gpu_check.py > gpurun_out/r2_dist_check_p$N.log 2>&1
  echo "dist check P=$N rc=$? ok=$(grep -c ' OK' gpurun_out/r2_dist_check_p$N.log) fail=$(grep -c 'FAIL' gpurun_out/r2_dist_check_p$N.log)"
  grep -v " OK" gpurun_out/r2_dist_check_p$N.log | grep -v "^\*\|OMP_NUM\|NCCL version\|^$" | tail -15
fi
$TR --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -1 gpurun_out/r2_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('P=%d: %.3f ms natural (%.0f GFLOP/s), transposed-out %.3f ms, check %s, stages %s, nvlink %s' % (d['n_gpus'], d['ms_per_step'], d['value'], d['config']['transposed_out_ms_per_step'], d.get('check'), d['roofline']['nvlink'].get('stage_ms'), {k: d['roofline']['nvlink'].get(k) for k in ('counter_tx_bytes_per_step','stage_exchange_gbs','sent_bytes_per_gpu_per_step')}))
print('e2e', d['e2e'])
" || tail -20 gpurun_out/r2_bench_n$N.err
