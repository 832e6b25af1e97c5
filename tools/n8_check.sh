#!/bin/bash
mkdir -p gpurun_out/r2z
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra > gpurun_out/r2z/n8_final.json 2> gpurun_out/r2z/n8_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z/n8_final.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], [round(x, 2) for x in d["roofline"]["kernel_ms_per_rank"]], d["parity_check"]["pass"], d["config"]["n_slabs"], d["clocks"]["sm_mhz"], round(d["roofline"]["frac"], 3))
PY
