for prec in f d; do for n in 16384 65536 262144 1048576 2097152; do for ip in 0 1; do python tools/dbg_case.py r2c $prec $n 3 $ip 2>&1 | tail -1; done; done; done
python tools/bench_configs.py --only C2 --flags estimate 2>&1 | cut -c1-600
python tools/r2r_experiment.py 2>&1 | cut -c1-420
python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -15
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "r2c or r2r or config or verifier" 2>&1 | tail -5
