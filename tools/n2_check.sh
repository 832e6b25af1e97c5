#!/bin/bash
mkdir -p gpurun_out/r2z
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2z/pytest_multi_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z/pytest_multi_final.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra > gpurun_out/r2z/n2_partial.json 2> gpurun_out/r2z/n2_partial.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2z/n2_partial.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), d["ms_per_step"], d["roofline"]["kernel_ms_per_rank"], d["parity_check"]["pass"], d["config"]["n_slabs"], d["clocks"]["sm_mhz"])
PY
