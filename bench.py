#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 retrieval engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): queries/sec at top-k=50 over a synthetic 10M x 1024 bf16 corpus, 4096-query batch.
One "step" = one pass of the hot path over one batch: fused similarity + top-k (tcgen05 GEMM with the top-k
epilogue, K1) + the per-query reduction of the slab lists; with N > 1 GPUs the corpus is sharded by node
range (strong scaling: the 10M rows are split, queries replicated) and the step adds the single candidate
all-gather + device k-way merge.

  value : whole-job queries/s with the query batch already resident in HBM.
  e2e   : same metric through the host-buffer C-ABI call (ahr_topk_host): pinned host queries -> H2D copy ->
          top-k -> D2H copy of scores/ids, all inside the timed region.
  roofline : the dominant kernel (K1) -- algorithmic flops 2*Q*N*D per launch over its CUDA-event duration
             on its launch stream, against the measured bf16 tensor peak in MEASURED_PEAKS.json.
  cpu_baseline : the oracle's CPU retrieval pass (oracle/cpu_retrieval.py) on a bounded sample, host cores.

``--impl reference`` times that CPU pass alone (the reference is pure Python over Chroma/HNSW, which is not
installable here) and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ROWS = int(os.environ.get("AHR_BENCH_ROWS", 10_000_000))
DIM = int(os.environ.get("AHR_BENCH_DIM", 1024))
N_QUERIES = int(os.environ.get("AHR_BENCH_QUERIES", 4096))
TOP_K = int(os.environ.get("AHR_BENCH_K", 50))
GEN_CHUNK = 250_000
METRIC = "queries/sec (top-k=50, 10M x 1024 bf16 corpus, 4096-query batch)"


def _traffic(world: int):
    """DRAM bytes per K1 launch from the committed ncu --set full capture (1-GPU c3 workload only)."""
    path = os.path.join(ROOT, "profiles", "r1_k1_traffic.json")
    if world == 1 and N_ROWS == 10_000_000 and DIM == 1024 and N_QUERIES == 4096 and os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops": float(p.get("bf16_tflops_sustained") or p.get("bf16_tflops")), "hbm_gbs": float(p["hbm_gbs"]),
                "tflops_burst": float(p.get("bf16_tflops") or 0.0),
                "source": "measured (MEASURED_PEAKS.json, bf16_tflops_sustained: the kernel is timed inside a long step)"}
    return {"tflops": 1400.0, "hbm_gbs": 6650.0, "tflops_burst": 1590.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up (nvidia-smi needs a few hundred ms to produce its first line) and only the samples
    whose timestamp falls inside the timed window are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def start(self) -> None:
        try:
            fd, self.path = tempfile.mkstemp(prefix="ahr_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def window_begin(self) -> None:
        self.t0 = time.time()

    def window_end(self) -> None:
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)  # let the last in-window sample land
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        import datetime
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            with open(self.path) as f:
                for line in f:
                    parts = [x.strip() for x in line.split(",")]
                    if len(parts) < 10:
                        continue
                    try:
                        ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                        rows.append((ts, float(parts[2]), float(parts[3]),
                                     [nm for nm, val in zip(names, parts[6:10]) if val.lower().startswith("active")]))
                    except ValueError:
                        continue
            os.remove(self.path)
        except Exception:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= (self.t1 or r[0]) + 0.05]
        scope = "timed window"
        if not inside:  # window shorter than one sampling period: report the samples under the same load around it
            inside, scope = rows[-5:], "last samples of the run (timed window shorter than the sampling period)"
        sm = [r[1] for r in inside]
        reasons = sorted({x for r in inside for x in r[3]})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "reasons": reasons, "samples": len(sm), "scope": scope}


def make_corpus_shard(torch, device, lo: int, hi: int, dim: int):
    """Rows [lo, hi) of the synthetic corpus: N(0,1) -> L2-normalise in fp32 -> bf16, generated in fixed
    250k-row chunks seeded by chunk index so every sharding sees the same 10M rows."""
    out = torch.empty((hi - lo, dim), dtype=torch.bfloat16, device=device)
    c0, c1 = lo // GEN_CHUNK, (hi + GEN_CHUNK - 1) // GEN_CHUNK
    for c in range(c0, c1):
        g = torch.Generator(device=device).manual_seed(1234 + c)
        rows = min(GEN_CHUNK, N_ROWS - c * GEN_CHUNK)
        x = torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32)
        x /= x.norm(dim=1, keepdim=True)
        a, b = max(lo, c * GEN_CHUNK), min(hi, c * GEN_CHUNK + rows)
        out[a - lo:b - lo] = x[a - c * GEN_CHUNK:b - c * GEN_CHUNK].to(torch.bfloat16)
        del x
    return out


def make_queries(torch, device, nq: int, dim: int, seed: int = 5678):
    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.randn((nq, dim), generator=g, device=device, dtype=torch.float32)
    q /= q.norm(dim=1, keepdim=True)
    return q.to(torch.bfloat16).contiguous()


def run_reference(args) -> int:
    """CPU arm: the oracle port of the reference's retrieval on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import cpu_retrieval
    sample_rows = int(os.environ.get("AHR_BENCH_CPU_ROWS", 500_000))
    sample_q = int(os.environ.get("AHR_BENCH_CPU_QUERIES", 128))
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_retrieval.time_sample(N_ROWS, DIM, TOP_K, 65536, 16)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_retrieval.time_sample(N_ROWS, DIM, TOP_K, sample_rows, sample_q))
    wall = time.perf_counter() - t0
    v = statistics.median([x["value"] for x in vals])
    base = vals[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / max(1, args.steps),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c3: synthetic {N_ROWS}x{DIM} corpus, {N_QUERIES}-query batch, top-k={TOP_K}",
                   "note": "reference CPU path = oracle port (torch fp32 GEMM + topk, all host threads); the "
                           "reference's own Chroma/HNSW call is not installable here"},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": base["cores"], "kind": "port", "sample": base["sample"]},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "exact"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    # stdout carries exactly ONE JSON line: everything else any library prints there (NCCL's version banner is a C-level
    # write to fd 1) is sent to stderr; the line itself goes to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        print("bench.py: no CUDA device -- the engine has no CPU fallback", file=sys.stderr)
        return 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the ONE JSON line, nothing else
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)

    from ah_rag_b200 import engine
    from ah_rag_b200.dist import ShardedIndex, shard_bounds

    lo, hi = shard_bounds(N_ROWS, world, rank)
    rows = make_corpus_shard(torch, device, lo, hi, DIM)
    sharded = ShardedIndex.from_rows(rows, id_base=lo)
    index = sharded.index
    exchange = "none (one shard)"
    if world > 1:
        # candidate exchange: stores into the peers' memory over NVLink (csrc/exchange.cu); AHR_EXCHANGE=nccl keeps the
        # all-gather.  The choice is made by every rank together (a failed IPC mapping on one rank falls back on all).
        if sharded.enable_peer_exchange(N_QUERIES, TOP_K):  # collective; True on all ranks or False on all ranks
            exchange = "peer-memory stores over NVLink + epoch flags (csrc/exchange.cu), device k-way merge"
        else:
            why = getattr(sharded, "peer_failure", "") or os.environ.get("AHR_EXCHANGE", "")
            if rank == 0 and why:
                print(f"bench.py: peer exchange not used ({why}); NCCL all-gather instead", file=sys.stderr)
            exchange = "single NCCL all-gather + device k-way merge"
    q_dev = make_queries(torch, device, N_QUERIES, DIM)
    q_host = q_dev.cpu().pin_memory()
    out_host = (torch.empty((N_QUERIES, TOP_K), dtype=torch.float32).pin_memory(),
                torch.empty((N_QUERIES, TOP_K), dtype=torch.int64).pin_memory())

    q_stage = torch.empty_like(q_dev)  # device staging of the end-to-end path (multi-GPU)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return sharded.topk(q_dev, TOP_K, mode=args.mode)

    def step_e2e():
        if world == 1:
            return index.topk_host(q_host, TOP_K, mode=args.mode, out=out_host)
        # multi-GPU: host queries -> device, sharded top-k (+ exchange + merge), result back to the host
        q_stage.copy_(q_host, non_blocking=True)
        s, i = sharded.topk(q_stage, TOP_K, mode=args.mode)
        out_host[0].copy_(s, non_blocking=True)
        out_host[1].copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)  # max over ranks
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step_device()
    step_e2e()

    sampler.window_begin()
    index.set_profiling(True)
    index.get_profile()
    launches0 = engine.launch_count()
    ms_total = timed(step_device, args.steps)
    launches = engine.launch_count() - launches0
    prof = index.get_profile()
    index.set_profiling(False)
    ms_e2e = timed(step_e2e, args.steps)
    sampler.window_end()
    clocks = sampler.stop() if rank == 0 else None

    ms_step = ms_total / args.steps
    qps = N_QUERIES / (ms_step / 1e3)
    qps_e2e = N_QUERIES / (ms_e2e / args.steps / 1e3)
    peaks = _peaks()
    k1_ms = prof["k1_ms_total"] / max(1, prof["k1_launches"])
    flops = 2.0 * N_QUERIES * (hi - lo) * DIM  # algorithmic flops of one K1 launch on this rank's shard
    achieved = flops / (k1_ms / 1e3) / 1e12 if k1_ms > 0 else 0.0

    if rank == 0:
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16" if args.mode == "bf16" else "f64-ordered", "data": "synthetic",
            "config": {"workload": f"c3: synthetic {N_ROWS}x{DIM} bf16 corpus (seeded N(0,1), L2-normalised), "
                                   f"{N_QUERIES}-query batch, top-k={TOP_K}, mode={args.mode}",
                       "sharding": f"node-range, {world} shard(s) of {hi - lo} rows; queries replicated; "
                                   + (exchange if world > 1 else "no collective"),
                       "l2": "inputs larger than L2 (corpus shard %.1f GB vs 126 MB)" % ((hi - lo) * DIM * 2 / 1e9),
                       "n_slabs": prof["n_slabs"]},
            "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": N_QUERIES * DIM * 2,
                    "d2h_bytes_per_step": N_QUERIES * TOP_K * 12},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": _traffic(world),
                         "kernel": "ahr::topk_tc2_kernel (K1, CTA-pair tcgen05 GEMM + top-k epilogue)",
                         "algorithmic_flops_per_launch": flops,
                         "frac_of_burst_peak": (achieved / peaks["tflops_burst"]) if peaks.get("tflops_burst") else None,
                         "kernel_ms": k1_ms, "kernel_share_of_step": k1_ms / ms_step if ms_step else None,
                         "peak_source": peaks["source"],
                         "hbm_floor_ms": (hi - lo) * DIM * 2 / (peaks["hbm_gbs"] * 1e9) * 1e3},
            "clocks": clocks,
        }
        if not args.skip_cpu_baseline and world == 1:
            from oracle import cpu_retrieval  # checker / reported baseline only, never the measured path
            line["cpu_baseline"] = {k: v for k, v in cpu_retrieval.time_sample(
                N_ROWS, DIM, TOP_K, int(os.environ.get("AHR_BENCH_CPU_ROWS", 500_000)),
                int(os.environ.get("AHR_BENCH_CPU_QUERIES", 128))).items() if k != "seconds"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
