#!/bin/sh
# Run the GPU tests one file per process (a crash in one file then cannot hide the others) and keep
# every log under gpurun_out/gputests/.
mkdir -p gpurun_out/gputests
for f in tests/test_*.py; do
    b=$(basename $f .py)
    timeout 1500 python -m pytest $f -m gpu -q -x > gpurun_out/gputests/$b.log 2>&1
    echo "$b rc=$? $(grep -E 'passed|failed|error|no tests ran|deselected' gpurun_out/gputests/$b.log | tail -1)"
done
