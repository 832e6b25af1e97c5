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

  parity_check : after the timed loops, at every N: sampled queries of the timed batch (bf16 mode: recall@k and score
             tolerance; fp32-exact mode: ids and score bits) against oracle/exact_topk.c scanning ALL corpus rows on
             the host (every rank scans its own shard, rank 0 merges with the same total order).
  extra    : driver-timed figures for the other BASELINE configurations -- c1 (the hotpotqa_mini graph through the drop-in
             Python API, B = 64 and B = 1, the reference's algorithm on one host core beside it), c2 (substitute: 450k x 384 fp32 index, 7405
             queries, fp32-exact top-5 + anchor expansion + rerank, one replica per GPU), c4 (100M x 768, 256-query batch: the real
             sharded configuration at N > 1, the per-GPU shard of the 8-GPU case at N = 1) and c5 (20M-node / 500M-edge
             graph, 1024 environments: hop 1, hop 2, uncapped-walk bandwidth, gym rollout) with the graph-half CPU
             baseline.  ``--no-extra`` skips it.

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
    """(DRAM bytes per K1 launch, where the figure comes from).  ncu cannot run inside the timed process, so the figure
    is the one of the committed ``ncu --set full`` capture of this same command (1-GPU c3 workload only)."""
    for name in ("r2_k1_traffic.json", "r1_k1_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if world == 1 and N_ROWS == 10_000_000 and DIM == 1024 and N_QUERIES == 4096 and os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            return d.get("dram_bytes_per_launch"), f"committed ncu capture profiles/{name} ({d.get('commit', 'commit not recorded')})"
    return None, "no ncu capture for this configuration (ncu runs one GPU, the c3 shape)"


def _host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


def _cpu_sample():
    """One bounded sample of the CPU retrieval pass: the REAL query batch against a slice of the corpus rows, all host
    threads (launchers such as torchrun export OMP_NUM_THREADS=1: set the thread count explicitly)."""
    import torch
    from oracle import cpu_retrieval
    torch.set_num_threads(_host_threads())
    return cpu_retrieval.time_sample(N_ROWS, DIM, TOP_K, int(os.environ.get("AHR_BENCH_CPU_ROWS", 250_000)),
                                     int(os.environ.get("AHR_BENCH_CPU_QUERIES", N_QUERIES)))


def run_reference(args) -> int:
    """CPU arm: the oracle port of the reference's retrieval on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from oracle import cpu_retrieval
    torch.set_num_threads(_host_threads())
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_retrieval.time_sample(N_ROWS, DIM, TOP_K, 65536, 64)
    vals = [_cpu_sample() for _ in range(args.steps)]
    v = statistics.median([x["value"] for x in vals])
    base = vals[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # time one full step (the whole batch against the whole corpus) takes at the measured rate
        "ms_per_step": 1000.0 * N_QUERIES / v,
        "sample_seconds_per_step": statistics.median([x["seconds"] for x in vals]),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"c3: synthetic {N_ROWS}x{DIM} corpus, {N_QUERIES}-query batch, top-k={TOP_K}",
                   "note": "reference CPU path = oracle port (torch fp32 GEMM + topk, all host threads); the "
                           "reference's own Chroma/HNSW call is not installable here"},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": base["cores"], "host_cpus": base["host_cpus"],
                         "kind": "port", "sample": base["sample"]},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def parity_check(torch, dist, world, rank, device, rows, lo, q_dev, sharded, k):
    """Sampled queries of the timed batch against the CPU oracle over ALL corpus rows.  Every rank scans its own shard
    on its share of the host cores (oracle/exact_topk.c, canonical fp64-ordered scores), rank 0 merges the lists with the
    same total order (score desc, id asc) and compares: bf16 mode -> recall@k and relative score error; fp32-exact mode
    -> ids and score bits."""
    import numpy as np
    from oracle import exact  # the checker, never the measured path
    nq = int(q_dev.shape[0])
    sel = sorted({0, 1, nq // 2, nq - 1})
    t0 = time.perf_counter()
    s_b, i_b = sharded.topk(q_dev, k, mode="bf16")
    s_b, i_b = s_b[sel].cpu().numpy(), i_b[sel].cpu().numpy()
    s_x, i_x = sharded.topk(q_dev, k, mode="exact")
    s_x, i_x = s_x[sel].cpu().numpy(), i_x[sel].cpu().numpy()
    bits = lambda t: t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
    qs = bits(q_dev[sel])
    threads = max(1, _host_threads() // world)
    ps, pi = [], []
    chunk = 1_000_000
    for c0 in range(0, int(rows.shape[0]), chunk):  # bounded host memory: 1M rows (2 GB) at a time
        sc, ids = exact.exact_topk_rowpar(bits(rows[c0:c0 + chunk]), qs, k, id_base=lo + c0, n_threads=threads)
        ps.append(sc)
        pi.append(ids)
    ls, li = exact.merge_topk(ps, pi, k)
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (ls, li))
        ls, li = exact.merge_topk([p[0] for p in parts], [p[1] for p in parts], k)
    recall = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(i_b, li)]))
    # bf16-mode scores against the canonical score of the same row (ids may differ at the k-th boundary)
    rel = 0.0
    for r in range(len(sel)):
        canon = dict(zip(li[r].tolist(), ls[r].tolist()))
        for idv, sv in zip(i_b[r].tolist(), s_b[r].tolist()):
            if idv in canon:
                rel = max(rel, abs(sv - canon[idv]) / max(abs(canon[idv]), 1e-6))
    ids_equal = bool(np.array_equal(i_x, li))
    bits_equal = bool(np.array_equal(s_x.view(np.uint32), ls.view(np.uint32)))
    return {"queries": sel, "rows_scanned": N_ROWS, "n_gpus": world,
            "exact_ids_equal": ids_equal, "exact_score_bits_equal": bits_equal,
            "bf16_recall_at_k": recall, "bf16_max_rel_score_err": rel,
            "pass": bool(ids_equal and bits_equal and recall >= 0.999 and rel <= 1e-2),
            "oracle": "oracle/exact_topk.c: full scan of every corpus row on the host (canonical fp64-ordered score, "
                      "ties -> lower id); per-rank shard scans merged with the same total order",
            "host_threads_per_rank": threads, "seconds": round(time.perf_counter() - t0, 2)}


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="bf16", choices=["bf16", "exact"])
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the c4 / c5 block")
    ap.add_argument("--no-parity", action="store_true", help="skip the sampled parity check against the CPU oracle")
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

    # N > 1, peer-memory exchange: back-to-back batches are pipelined -- batch i is exchanged and merged on a side stream
    # while K1 of batch i + 1 runs (dist.ShardedIndex.topk_async; at most two batches in flight, results consumed in order).
    # Without it every step costs the maximum over the ranks of that step's K1 (profiles/r2_scaling.md).
    pipelined = (world > 1 and getattr(sharded, "peer", None) is not None and args.mode == "bf16"
                 and os.environ.get("AHR_BENCH_PIPELINE", "1") != "0")
    in_flight = []

    def step_device():
        if not pipelined:
            return sharded.topk(q_dev, TOP_K, mode=args.mode)
        in_flight.append(sharded.topk_async(q_dev, TOP_K, mode=args.mode))
        return in_flight.pop(0).result() if len(in_flight) > 2 else None

    def drain():
        while in_flight:
            in_flight.pop(0).result()

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
        drain()  # (pipelined device loop: the last two batches' merges; a no-op otherwise)
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
    drain()
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

    parity = None
    if not args.no_parity:
        parity = parity_check(torch, dist, world, rank, device, rows, lo, q_dev, sharded, TOP_K)
    extra = None
    if not args.no_extra and os.environ.get("AHR_BENCH_EXTRA", "1") != "0":
        # free the c3 shard first: the c4 shard of a 2-GPU run is 76.8 GB
        del sharded, index, rows
        torch.cuda.empty_cache()
        import bench_workloads as bw
        extra = {}
        for name, fn in (("c1", bw.extra_c1), ("c2", bw.extra_c2), ("c4", bw.extra_c4), ("c5", bw.extra_c5)):
            try:  # a failing side block must not take the headline line with it; the error is reported in its place
                extra[name] = fn(dist, world, rank, device)
            except Exception as e:  # noqa: BLE001
                extra[name] = {"error": repr(e)}
                torch.cuda.empty_cache()

    ms_step = ms_total / args.steps
    qps = N_QUERIES / (ms_step / 1e3)
    qps_e2e = N_QUERIES / (ms_e2e / args.steps / 1e3)
    peaks = _peaks()
    k1_ms = prof["k1_ms_total"] / max(1, prof["k1_launches"])
    k1_all = [k1_ms]
    if world > 1:  # every rank's K1 time: the step waits for the slowest GPU (clocks differ under the power cap)
        t = torch.tensor([k1_ms], device=device, dtype=torch.float64)
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        k1_all = [float(x.item()) for x in g]
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
                                   + (exchange if world > 1 else "no collective")
                                   + ("; device loop pipelined: batch i exchanged and merged on a side stream while K1 of "
                                      "batch i + 1 runs (AHR_BENCH_PIPELINE=0: one stream)" if pipelined else ""),
                       "l2": "inputs larger than L2 (corpus shard %.1f GB vs 126 MB)" % ((hi - lo) * DIM * 2 / 1e9),
                       "n_slabs": prof["n_slabs"]},
            "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": N_QUERIES * DIM * 2,
                    "d2h_bytes_per_step": N_QUERIES * TOP_K * 12},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["tflops"], "traffic": _traffic(world)[0],
                         "traffic_source": _traffic(world)[1],
                         "kernel": "ahr::topk_tc2_kernel (K1, CTA-pair tcgen05 GEMM + top-k epilogue)",
                         "algorithmic_flops_per_launch": flops,
                         "frac_of_burst_peak": (achieved / peaks["tflops_burst"]) if peaks.get("tflops_burst") else None,
                         "kernel_ms": k1_ms, "kernel_share_of_step": k1_ms / ms_step if ms_step else None,
                         "kernel_ms_per_rank": k1_all,
                         "slowest_rank_kernel_share_of_step": max(k1_all) / ms_step if ms_step else None,
                         "peak_source": peaks["source"],
                         "hbm_floor_ms": (hi - lo) * DIM * 2 / (peaks["hbm_gbs"] * 1e9) * 1e3},
            "clocks": clocks,
            "timing_note": "value and e2e are two separate K-step loops run back to back; under the power cap the SM clock "
                           "drifts by ~2 % between them, so e2e can read slightly above value",
        }
        if parity is not None:
            line["parity_check"] = parity
        if extra is not None:
            line["extra"] = extra
        if not args.skip_cpu_baseline and world == 1:
            # checker / reported baseline only, never the measured path
            line["cpu_baseline"] = {k: v for k, v in _cpu_sample().items() if k != "seconds"}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
