#!/bin/bash
# same-box A/B of the partial bound on c3 (AHR_TC_PARTIAL_BOUND=0/1), with wait counters, K1 tests and the parity check
for o in 0 1 0 1; do
  echo "=== AHR_TC_PARTIAL_BOUND=$o"
  AHR_TC_PARTIAL_BOUND=$o python bench.py --steps 5 --warmup 3 --no-extra --skip-cpu-baseline --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' c3 qps', round(d['value']), 'k1_ms', round(d['roofline']['kernel_ms'],2), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
done
for o in 0 1; do
  AHR_TC_STATS=1 AHR_TC_PARTIAL_BOUND=$o python bench.py --steps 2 --warmup 3 --no-extra --skip-cpu-baseline --no-parity 2>&1 | grep leaders | tail -1 | cut -c1-190
done
python -m pytest tests/test_gpu_topk.py tests/test_gpu_scale.py -x -q 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-extra --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' parity', d['parity_check'])"
