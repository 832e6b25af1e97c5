#!/bin/bash
# A/B helper for gpurun: prints one summary line per (variant, prefetch, slabs) bench configuration
for cfg in "$@"; do
  set -- $cfg
  echo "=== variant $1 prefetch $2 slabs ${3:-auto}"
  AHR_TC_VARIANT=$1 AHR_TC_SLABS=${3:-0} timeout 300 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > /tmp/ab.json 2>/tmp/ab.err
  if [ -s /tmp/ab.json ]; then
    tail -1 /tmp/ab.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('qps',round(d['value']), 'ms',round(d['ms_per_step'],2), 'TF',round(d['roofline']['achieved'],1), 'frac',round(d['roofline']['frac'],3), 'e2e',round(d['e2e']['value']), 'slabs',d['config']['n_slabs'], d['clocks'])"
  else
    echo "FAILED"; tail -5 /tmp/ab.err
  fi
done
