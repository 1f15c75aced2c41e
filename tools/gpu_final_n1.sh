#!/bin/sh
# final single-GPU evidence of a round: every GPU test, smoke, the bench line, its ncu launch list
set -x
sh tools/run_gpu_tests_by_file.sh > gpurun_out/r02_gputests_summary.log 2>&1
cat gpurun_out/r02_gputests_summary.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --wisdom gpurun_out/r02_wisdom.txt > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 400 gpurun_out/r2_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra --no-check --wisdom gpurun_out/r02_wisdom.txt > gpurun_out/r2_bench_ncu.json 2> gpurun_out/r2_bench_ncu.err
wc -l gpurun_out/r02_ncu_launches_bench.csv
