set -x
python bench.py --steps 10 --warmup 3 --wisdom gpurun_out/r02_wisdom.txt > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
tail -c 300 gpurun_out/r2_bench_c.json
# launch list of one bench step (wisdom imported: no planner launches); shares of the step per kernel
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extra --no-check --wisdom gpurun_out/r02_wisdom.txt > gpurun_out/r2_bench_ncu.json 2> gpurun_out/r2_bench_ncu.err
wc -l gpurun_out/r02_ncu_launches_bench.csv
mkdir -p /tmp/prof
ncu --set full --clock-control none -k regex:warp1024 -s 2 -c 1 -o /tmp/prof/rows_warp python tools/one_pass.py 54 2 256 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fast_kernel -s 2 -c 1 -o /tmp/prof/rows_block python tools/one_pass.py 13 2 256 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fast_kernel -s 2 -c 1 -o /tmp/prof/dim1_tile4 python tools/one_pass.py 50 1 256 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fast_kernel -s 2 -c 1 -o /tmp/prof/dim0_tile8 python tools/one_pass.py 15 0 1024 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fast_kernel -s 2 -c 2 -o /tmp/prof/c2_r2c python tools/c2_pass.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fast_kernel -s 2 -c 1 -o /tmp/prof/c5b_r2r python tools/r2r_pass.py > /dev/null 2>&1
ls -la /tmp/prof
python tools/ncu_summary.py gpurun_out/r02_ncu_full_summary.txt /tmp/prof/*.ncu-rep
cat gpurun_out/r02_ncu_full_summary.txt | head -150
