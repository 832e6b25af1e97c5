#!/bin/bash
# same-box A/B of the K1 list filter: c4 shard (Q = 256 and the batch-size sweep) and c3, filter off / on / off / on
for f in 0 1 0 1; do
  echo "=== AHR_TC_LIST_FILTER=$f"
  AHR_TC_LIST_FILTER=$f python bench_workloads.py c4 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 Q=256 k1_ms', round(d['k1_ms'],3), 'tensor_frac', round(d['roofline']['tensor_frac'],3))"
done
for f in 0 1; do
  echo "=== AHR_TC_LIST_FILTER=$f"
  AHR_TC_LIST_FILTER=$f python bench_workloads.py qsweep 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d.get('variant')==2 or d['nq']<=128: print(' nq',d['nq'],'variant',d['variant'],'k1_ms',d['k1_ms'])"
  AHR_TC_LIST_FILTER=$f python bench.py --steps 5 --warmup 3 --no-extra --skip-cpu-baseline --no-parity 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(' c3 qps', round(d['value']), 'k1_ms', round(d['roofline']['kernel_ms'],2), d['clocks']['sm_mhz'])"
done
