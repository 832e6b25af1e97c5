#!/bin/bash
# wait counters (cycles: clock independent) of K1 on the c4 shard at Q = 512 / 1024 and on c3 for the three filter settings
for f in 0 1 2; do
  echo "=== AHR_TC_LIST_FILTER=$f"
  AHR_TC_STATS=1 AHR_TC_LIST_FILTER=$f AHR_SWEEP_NQ=512,1024 python bench_workloads.py qsweep 2>&1 | grep -E "variant\": 2|leaders" | awk '/leaders/{l=$0} /variant/{print l; print $0}' | cut -c1-230
  AHR_TC_STATS=1 AHR_TC_LIST_FILTER=$f python bench.py --steps 3 --warmup 3 --no-extra --skip-cpu-baseline --no-parity 2>&1 | grep -E "leaders" | tail -1 | cut -c1-200
done
