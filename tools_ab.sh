#!/bin/bash
# A/B helper for gpurun: one summary line of bench.py per "variant slabs" configuration, e.g.
#   gpurun -- 'bash tools_ab.sh "0 0" "2 9" "2 12"'        (variant 0 = by batch size, 1 = single CTA, 2 = CTA pairs; slabs 0 = auto)
for cfg in "$@"; do
  set -- $cfg
  echo "=== variant $1 slabs ${2:-auto}"
  AHR_TC_VARIANT=$1 AHR_TC_SLABS=${2:-0} timeout 300 python bench.py --steps 3 --warmup 3 --skip-cpu-baseline > /tmp/ab.json 2>/tmp/ab.err
  if [ -s /tmp/ab.json ]; then
    tail -1 /tmp/ab.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('qps',round(d['value']), 'ms',round(d['ms_per_step'],2), 'TF',round(d['roofline']['achieved'],1), 'frac',round(d['roofline']['frac'],3), 'e2e',round(d['e2e']['value']), 'slabs',d['config']['n_slabs'], d['clocks'])"
  else
    echo "FAILED"; tail -5 /tmp/ab.err
  fi
done
