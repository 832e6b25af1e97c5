#!/bin/bash
# one iteration on the wide-expand kernel: parity tests, the c5 workload, then an ncu capture of the bandwidth view
tag=${1:-x}
mkdir -p gpurun_out/r2r
python -m pytest tests/test_gpu_graph.py -x -q -k "large_limits" > gpurun_out/r2r/pytest_$tag.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/r2r/pytest_$tag.log
python bench_workloads.py c5 > gpurun_out/r2r/c5_$tag.log 2>&1; echo c5 rc=$?
tail -1 gpurun_out/r2r/c5_$tag.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:v for k,v in d.items() if k.startswith('hop') or 'exact' in k})"
ncu --set full --clock-control none --import-source on -k regex:expand_wide_kernel -s 24 -c 1 -o gpurun_out/r2r/k2_wide_$tag python bench_workloads.py c5 > gpurun_out/r2r/c5_ncu_$tag.log 2>&1; echo ncu rc=$?
