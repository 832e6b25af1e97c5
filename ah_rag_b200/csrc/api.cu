// api.cu -- the C ABI of libahrag_b200 (include/ahrag_b200.h): handles, workspaces, orchestration.
#include <cmath>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "tc_shared.cuh"

namespace ahr {

static thread_local std::string g_error;
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
}

}  // namespace ahr

using namespace ahr;

struct ahr_index {
  int device = 0;
  int sm_count = 148;
  int64_t n_rows = 0;
  int32_t dim = 0;
  int dtype = 0;
  const void *rows = nullptr;
  int64_t id_base = 0;
  bool has_map = false;
  CUtensorMap map_x;   // box 64 x 256 rows (single-CTA K1)
  CUtensorMap map_x2;  // box 64 x 128 rows (CTA-pair K1: each CTA stages half a corpus tile)
  int variant = 0;     // AHR_TC_VARIANT: 0 = by batch size (default), 1 = single CTA where it applies, 2 = CTA pairs
  int seed_stride = 0;   // AHR_TC_SEED_STRIDE: sample pass over every n-th corpus tile seeds the thresholds.  Off by
                         // default: exact and tested, but since the v5 epilogue it buys nothing measurable (history file)
  int force_slabs = 0;     // AHR_TC_SLABS: override the slab heuristic (experiments)
  bool shared_bound = true;  // AHR_TC_SHARED_BOUND=0 switches the cross-slab admission bound off (A/B runs)
  DevBuf x_hi, x_lo;   // fp32 index: bf16 hi/lo halves of the rows (library-owned), maps below point at them
  CUtensorMap map_x2_lo;
  DevBuf q_split;      // fp32 queries split per call
  bool tc_stats = false;   // AHR_TC_STATS=1: per-role wait-cycle counters of K1, printed to stderr (debug)
  DevBuf stats_buf;
  DevBuf ws;        // partial lists / score chunks
  DevBuf ws2;       // exact-mode scratch
  DevBuf hq, hs, hi;  // device staging of the *_host entry point
  float *d_max_norm = nullptr;
  int64_t n_certified = 0, n_fallback = 0;
  struct Pending {       // an exact query between ahr_topk_exact_begin and ahr_topk_exact_finish
    bool active = false, done = false, use_tc = false;
    const void *q = nullptr;
    int32_t nq = 0, k = 0;
    float *scores = nullptr;
    int64_t *ids = nullptr;
    cudaStream_t stream = nullptr;
  } pending;
  int32_t *h_nfail = nullptr;      // pinned: number of queries tier 1 could not certify
  cudaEvent_t ev_nfail = nullptr;  // recorded behind the copy into h_nfail
  bool profile = false;  // CUDA-event timing of the dominant kernel (K1) on its launch stream
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  int last_slabs = 0;
};

struct ahr_graph {
  int device = 0;
  GraphView view;
  DevBuf ws;
  DevBuf recs;  // packed node lines (one 64-byte NodeLine per node)
};

extern "C" {

const char *ahr_last_error(void) { return g_error.c_str(); }
int ahr_version(void) { return AHR_VERSION; }
int64_t ahr_launch_count(void) { return g_launches.load(); }

// ---------------------------------------------------------------------------------------------
int ahr_index_create(ahr_index **out, int device, int64_t n_rows, int32_t dim, int dtype, const void *dev_rows,
                     int64_t id_base) {
  AHR_REQUIRE(out != nullptr, "ahr_index_create: out is NULL");
  AHR_REQUIRE(n_rows >= 0 && dim >= 1, "ahr_index_create: bad shape [%lld, %d]", (long long)n_rows, dim);
  AHR_REQUIRE(dtype == AHR_DTYPE_F32 || dtype == AHR_DTYPE_BF16, "ahr_index_create: unknown dtype %d", dtype);
  AHR_REQUIRE(n_rows == 0 || dev_rows != nullptr, "ahr_index_create: dev_rows is NULL");
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("ahr_index_create: cannot select CUDA device %d: %s", device, cudaGetErrorString(cudaGetLastError()));
    return AHR_E_CUDA;
  }
  cudaDeviceProp prop;
  AHR_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("libahrag_b200 is built for sm_100a (Blackwell B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    return AHR_E_UNSUPPORTED;
  }
  ahr_index *idx = new ahr_index();
  idx->device = device;
  idx->sm_count = prop.multiProcessorCount;
  idx->n_rows = n_rows;
  idx->dim = dim;
  idx->dtype = dtype;
  idx->rows = dev_rows;
  idx->id_base = id_base;
  if (dtype == AHR_DTYPE_BF16 && n_rows > 0 && dim % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(dev_rows) & 15) == 0) {
    int rc = tc_make_corpus_map(&idx->map_x, dev_rows, n_rows, dim);
    if (rc == AHR_OK) rc = tc2_make_corpus_map(&idx->map_x2, dev_rows, n_rows, dim);
    if (rc != AHR_OK) { delete idx; return rc; }
    idx->has_map = true;
  }
  if (dtype == AHR_DTYPE_F32 && n_rows > 0 && dim % 8 == 0) {
    // tensor-core candidate generation for fp32 rows: keep hi/lo bf16 halves (same bytes as the fp32 original)
    const int64_t ne = n_rows * (int64_t)dim;
    int rc = idx->x_hi.reserve((size_t)ne * 2);
    if (rc == AHR_OK) rc = idx->x_lo.reserve((size_t)ne * 2);
    if (rc == AHR_OK) rc = split_f32_hi_lo(reinterpret_cast<const float *>(dev_rows), ne, idx->x_hi.ptr, idx->x_lo.ptr, 0);
    if (rc == AHR_OK) rc = tc2_make_corpus_map(&idx->map_x2, idx->x_hi.ptr, n_rows, dim);
    if (rc == AHR_OK) rc = tc2_make_corpus_map(&idx->map_x2_lo, idx->x_lo.ptr, n_rows, dim);
    if (rc != AHR_OK) { idx->x_hi.release(); idx->x_lo.release(); delete idx; return rc; }
    idx->has_map = true;
  }
  if (const char *e = getenv("AHR_TC_VARIANT")) idx->variant = std::min(2, std::max(0, atoi(e)));
  if (const char *e = getenv("AHR_TC_SEED_STRIDE")) idx->seed_stride = std::max(0, atoi(e));
  if (const char *e = getenv("AHR_TC_SLABS")) idx->force_slabs = std::max(0, atoi(e));
  if (const char *e = getenv("AHR_TC_STATS")) idx->tc_stats = atoi(e) != 0;
  if (const char *e = getenv("AHR_TC_SHARED_BOUND")) idx->shared_bound = atoi(e) != 0;
  cudaError_t e = cudaMalloc(&idx->d_max_norm, sizeof(float));
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void **>(&idx->h_nfail), sizeof(int32_t), cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&idx->ev_nfail, cudaEventDisableTiming);
  if (e != cudaSuccess) { delete idx; set_error("allocation failed: %s", cudaGetErrorString(e)); return AHR_E_NOMEM; }
  int rc = row_norms_max(dev_rows, dtype, n_rows, dim, idx->d_max_norm, 0);
  if (rc != AHR_OK) { cudaFree(idx->d_max_norm); delete idx; return rc; }
  AHR_CUDA_CHECK(cudaStreamSynchronize(0));
  *out = idx;
  return AHR_OK;
}

int ahr_index_destroy(ahr_index *idx) {
  if (!idx) return AHR_OK;
  DeviceGuard guard(idx->device);
  idx->x_hi.release(); idx->x_lo.release(); idx->q_split.release(); idx->stats_buf.release(); idx->ws.release(); idx->ws2.release(); idx->hq.release(); idx->hs.release(); idx->hi.release();
  if (idx->d_max_norm) cudaFree(idx->d_max_norm);
  if (idx->h_nfail) cudaFreeHost(idx->h_nfail);
  if (idx->ev_nfail) cudaEventDestroy(idx->ev_nfail);
  delete idx;
  return AHR_OK;
}

// brute-force exact scan for a subset of queries (q_sel == nullptr: all), results into out_* [nq_sel, k]
static int exact_bruteforce(ahr_index *idx, const void *dev_q, const int32_t *q_sel, int32_t nq_sel, int32_t k,
                            float *out_scores, int64_t *out_ids, DevBuf &scratch, cudaStream_t st) {
  if (nq_sel <= 0) return AHR_OK;
  int64_t chunk = (int64_t)(64ll << 20) / nq_sel;  // <= 256 MB of scores
  chunk = std::max<int64_t>(64, std::min<int64_t>(chunk, std::max<int64_t>(idx->n_rows, 1)));
  int rc = scratch.reserve((size_t)nq_sel * chunk * sizeof(float));
  if (rc != AHR_OK) return rc;
  float *sc = reinterpret_cast<float *>(scratch.ptr);
  bool first = true;
  if (idx->n_rows == 0) return select_topk(sc, nullptr, nullptr, 0, 0, nq_sel, k, 0, false, out_scores, out_ids, st);
  for (int64_t r0 = 0; r0 < idx->n_rows; r0 += chunk) {
    const int64_t n = std::min(chunk, idx->n_rows - r0);
    rc = exact_scores(idx->rows, dev_q, idx->dtype, idx->dim, r0, n, q_sel, nq_sel, sc, st);
    if (rc != AHR_OK) return rc;
    rc = select_topk(sc, nullptr, nullptr, n, n, nq_sel, k, idx->id_base + r0, !first, out_scores, out_ids, st);
    if (rc != AHR_OK) return rc;
    first = false;
  }
  return AHR_OK;
}

// tensor-core path: per-slab partial lists (K1) + per-query reduction, results [nq, k] with global ids
//
// Dispatch by batch size (profiles/r2_qsweep.md): up to 128 queries one CTA per tile (M = 128: the MMA takes half the
// time of the HBM stream, the kernel runs at the memory roofline); above, CTA pairs (M = 256 per corpus tile read).
// Queries are padded to a whole tile in a library buffer: TMA boxes that are mostly out of bounds are slow to fill
// (Q = 1 took 2x the time of Q = 128 on the same corpus).
// Optional sample pass (AHR_TC_SEED_STRIDE=n): a 1/n sample of the corpus tiles is scored first and its k-th best score
// per query seeds every list's admission threshold in the main pass (tc_shared.cuh: seed_threshold).
static int tc_topk(ahr_index *idx, const void *dev_q, int32_t nq, int32_t k, float *out_scores, int64_t *out_ids,
                   float *dbg_scores, cudaStream_t st, const PeerPush *push = nullptr) {
  AHR_NVTX("ahr K1 fused similarity + top-k");
  AHR_REQUIRE(idx->has_map, "index is not TMA-addressable (dim %% 8 != 0 or rows not 16-byte aligned)");
  AHR_REQUIRE((reinterpret_cast<uintptr_t>(dev_q) & 15) == 0, "queries must be 16-byte aligned");
  const bool split = idx->dtype == AHR_DTYPE_F32;  // fp32 operands: hi + lo bf16 halves, 3 operand segments
  const bool pairs = split || k > 56 || idx->variant == 2 || (idx->variant == 0 && nq > 128);
  const int32_t q_tile = pairs ? 256 : 128;
  const int32_t nq_pad = (nq + q_tile - 1) / q_tile * q_tile;
  const void *q_hi = dev_q, *q_lo = nullptr;
  const int64_t ne = (int64_t)nq * idx->dim, ne_pad = (int64_t)nq_pad * idx->dim;
  if (split) {
    int rc0 = idx->q_split.reserve((size_t)ne_pad * 4 + 512);
    if (rc0 != AHR_OK) return rc0;
    uint8_t *base = reinterpret_cast<uint8_t *>(idx->q_split.ptr);
    void *hi = base, *lo = base + (((size_t)ne_pad * 2 + 255) & ~(size_t)255);
    rc0 = split_f32_hi_lo(reinterpret_cast<const float *>(dev_q), ne, hi, lo, st);
    if (rc0 != AHR_OK) return rc0;
    if (nq_pad > nq) {
      AHR_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<uint8_t *>(hi) + ne * 2, 0, (size_t)(ne_pad - ne) * 2, st));
      AHR_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<uint8_t *>(lo) + ne * 2, 0, (size_t)(ne_pad - ne) * 2, st));
    }
    q_hi = hi;
    q_lo = lo;
  } else if (nq_pad > nq) {
    int rc0 = idx->q_split.reserve((size_t)ne_pad * 2);
    if (rc0 != AHR_OK) return rc0;
    AHR_CUDA_CHECK(cudaMemcpyAsync(idx->q_split.ptr, dev_q, (size_t)ne * 2, cudaMemcpyDeviceToDevice, st));
    AHR_CUDA_CHECK(cudaMemsetAsync(reinterpret_cast<uint8_t *>(idx->q_split.ptr) + ne * 2, 0, (size_t)(ne_pad - ne) * 2, st));
    q_hi = idx->q_split.ptr;
  }
  const int64_t n_ctiles = (idx->n_rows + 255) / 256;
  const int32_t n_qtiles = nq_pad / q_tile;
  const int n_units = pairs ? std::max(1, idx->sm_count / 2) : idx->sm_count;  // CTA pairs or CTAs
  int S = tc_choose_slabs(n_qtiles, n_ctiles, n_units);
  if (idx->force_slabs > 0) S = (int)std::min<int64_t>(idx->force_slabs, n_ctiles);
  // sample pass geometry: every seed_stride-th corpus tile; worth it only when the sample is a few thousand rows
  const int stride = (!dbg_scores && idx->seed_stride > 1 && n_ctiles / idx->seed_stride >= 32) ? idx->seed_stride : 1;
  const int64_t n_stiles = (n_ctiles + stride - 1) / stride;
  const int Ss = stride > 1 ? tc_choose_slabs(n_qtiles, n_stiles, n_units) : 0;
  const size_t n_part = (size_t)nq * S * k, n_spart = (size_t)nq * Ss * k;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_ps = carve(n_part * 4), o_pi = carve(n_part * 4), o_ss = carve(n_spart * 4), o_si = carve(n_spart * 4),
               o_ts = carve((size_t)nq * k * 4), o_ti = carve((size_t)nq * k * 8), o_tb = carve((size_t)nq_pad * 4 * std::max(S, Ss) * 2);  // full + partial bound
  int rc = idx->ws.reserve(off);
  if (rc != AHR_OK) return rc;
  uint8_t *w = reinterpret_cast<uint8_t *>(idx->ws.ptr);
  float *ps = reinterpret_cast<float *>(w + o_ps);
  int32_t *pi = reinterpret_cast<int32_t *>(w + o_pi);
  idx->last_slabs = S;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (idx->profile) {
    AHR_CUDA_CHECK(cudaEventCreate(&e0));
    AHR_CUDA_CHECK(cudaEventCreate(&e1));
    AHR_CUDA_CHECK(cudaEventRecord(e0, st));
  }
  unsigned long long *stats = nullptr;
  if (idx->tc_stats && pairs) {
    rc = idx->stats_buf.reserve((size_t)idx->sm_count * 8 * 8);
    if (rc != AHR_OK) return rc;
    stats = reinterpret_cast<unsigned long long *>(idx->stats_buf.ptr);
    AHR_CUDA_CHECK(cudaMemsetAsync(stats, 0, (size_t)idx->sm_count * 64, st));
  }
  auto launch = [&](int slabs, float *part_s, int32_t *part_i, const TcExtra &ex, unsigned long long *stats_p) {
    if (pairs)
      return tc2_topk_partials(idx->map_x2, split ? &idx->map_x2_lo : nullptr, q_hi, q_lo, idx->n_rows, idx->dim, nq, k,
                               slabs, n_units, part_s, part_i, dbg_scores, stats_p, ex, st);
    return tc_topk_partials(idx->map_x, q_hi, idx->n_rows, idx->dim, nq, k, slabs, idx->sm_count, part_s, part_i,
                            dbg_scores, ex, st);
  };
  TcExtra ex;
  ex.q_rows = nq_pad;
  if (idx->shared_bound && !dbg_scores) {
    // what every slab publishes for every query (tc_shared.cuh, "shared bound"): 0 = nothing published yet
    ex.thr_shared = reinterpret_cast<uint32_t *>(w + o_tb);
    AHR_CUDA_CHECK(cudaMemsetAsync(ex.thr_shared, 0, (size_t)nq_pad * 4 * S * 2, st));
  }
  if (stride > 1) {
    float *ss = reinterpret_cast<float *>(w + o_ss), *ts = reinterpret_cast<float *>(w + o_ts);
    int32_t *si = reinterpret_cast<int32_t *>(w + o_si);
    int64_t *ti = reinterpret_cast<int64_t *>(w + o_ti);
    TcExtra exs = ex;
    exs.tile_stride = stride;
    exs.thr_shared = nullptr;  // the sample pass is short and keeps plain lists
    rc = launch(Ss, ss, si, exs, nullptr);
    if (rc != AHR_OK) return rc;
    rc = select_topk(ss, si, nullptr, (int64_t)Ss * k, (int64_t)Ss * k, nq, k, 0, false, ts, ti, st);
    if (rc != AHR_OK) return rc;
    ex.thr_src = ts + (k - 1);  // k-th best sample score of each query (-inf when the sample holds fewer than k rows)
    ex.thr_stride = k;
  }
  rc = launch(S, ps, pi, ex, stats);
  if (rc != AHR_OK) return rc;
  if (idx->profile) {
    AHR_CUDA_CHECK(cudaEventRecord(e1, st));
    idx->prof_events.emplace_back(e0, e1);
  }
  if (stats) {
    std::vector<unsigned long long> h((size_t)idx->sm_count * 8);
    AHR_CUDA_CHECK(cudaMemcpyAsync(h.data(), stats, h.size() * 8, cudaMemcpyDeviceToHost, st));
    AHR_CUDA_CHECK(cudaStreamSynchronize(st));
    double sum[8] = {0}, total = 0;
    int n_lead = 0;
    for (int b = 0; b < idx->sm_count; ++b) {
      for (int j = 0; j < 8; ++j) sum[j] += (double)h[(size_t)b * 8 + j];
      if (h[(size_t)b * 8 + 3]) { total += (double)h[(size_t)b * 8 + 3]; ++n_lead; }
    }
    if (n_lead)
      fprintf(stderr, "[ahr tc stats] leaders=%d avg cycles: total=%.0f mma_wait_full=%.0f mma_wait_tempty=%.0f | "
              "per CTA: producer_wait_empty=%.0f epi_wait_tfull=%.0f epi_slow_chunks=%.0f epi_slow_cycles=%.0f epi_row_compactions=%.0f\n", n_lead,
              total / n_lead, sum[1] / n_lead, sum[2] / n_lead, sum[0] / idx->sm_count, sum[4] / idx->sm_count,
              sum[5] / idx->sm_count, sum[6] / idx->sm_count, sum[7] / idx->sm_count);
  }
  AHR_NVTX("ahr K1 slab-list reduction (+ peer push)");
  // per-query reduction of the S slab lists (each sorted, empties last): rank-counting merge for a few lists, bitonic
  // selection otherwise; with `push` the result also goes to the peers of the sharded index from the same launch
  if (merge_sorted_applicable(S, k, k))
    return merge_sorted_lists(ps, pi, nullptr, (int64_t)S * k, S, k, k, k, nq, k, idx->id_base, out_scores, out_ids, push, st);
  rc = select_topk(ps, pi, nullptr, (int64_t)S * k, (int64_t)S * k, nq, k, idx->id_base, false, out_scores, out_ids, st);
  if (rc == AHR_OK && push != nullptr && push->world > 1) rc = peer_push_record(*push, (size_t)nq * k * 12, st);
  return rc;
}

}  // extern "C"

namespace ahr {
int index_topk_bf16(ahr_index *idx, const void *dev_q, int32_t nq, int32_t k, float *dev_scores, int64_t *dev_ids,
                    const PeerPush *push, cudaStream_t stream) {
  AHR_REQUIRE(idx != nullptr && dev_q && dev_scores && dev_ids, "ahr_topk: NULL argument");
  AHR_REQUIRE(nq >= 1 && k >= 1 && k <= 120, "AHR_MODE_BF16 supports 1 <= k <= 120 (got %d)", k);
  DeviceGuard guard(idx->device);
  if (idx->n_rows == 0) {
    int rc = select_topk(dev_scores, nullptr, nullptr, 0, 0, nq, k, 0, false, dev_scores, dev_ids, stream);
    if (rc == AHR_OK && push != nullptr && push->world > 1) rc = peer_push_record(*push, (size_t)nq * k * 12, stream);
    return rc;
  }
  return tc_topk(idx, dev_q, nq, k, dev_scores, dev_ids, nullptr, stream, push);
}
}  // namespace ahr

extern "C" {

int ahr_topk(ahr_index *idx, const void *dev_q, int32_t nq, int32_t k, int mode, float *dev_scores, int64_t *dev_ids,
             void *stream) {
  AHR_REQUIRE(idx != nullptr, "ahr_topk: index is NULL");
  AHR_REQUIRE(nq >= 0 && k >= 1, "ahr_topk: bad sizes nq=%d k=%d", nq, k);
  AHR_REQUIRE(k <= SELECT_MAX_K, "ahr_topk: k must be <= %d (got %d)", SELECT_MAX_K, k);
  AHR_REQUIRE(mode == AHR_MODE_BF16 || mode == AHR_MODE_EXACT, "ahr_topk: unknown mode %d", mode);
  if (nq == 0) return AHR_OK;
  AHR_REQUIRE(dev_q && dev_scores && dev_ids, "ahr_topk: NULL buffer");
  DeviceGuard guard(idx->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

  if (mode == AHR_MODE_BF16) {
    AHR_REQUIRE(k <= 120, "AHR_MODE_BF16 supports k <= 120 (got %d); use AHR_MODE_EXACT", k);
    if (idx->n_rows == 0) return select_topk(dev_scores, nullptr, nullptr, 0, 0, nq, k, 0, false, dev_scores, dev_ids, st);
    return tc_topk(idx, dev_q, nq, k, dev_scores, dev_ids, nullptr, st);
  }

  int rc = ahr_topk_exact_begin(idx, dev_q, nq, k, dev_scores, dev_ids, stream);
  if (rc != AHR_OK) return rc;
  return ahr_topk_exact_finish(idx);
}

// ---- exact mode ---------------------------------------------------------------------------------------------------
// Tiers: (1) tensor-core candidates with a small slack (kc = 56 for k <= 50: the fast 64-slot lists), canonical rescoring,
// certification; (2) the queries that could not be certified (near-ties, duplicated rows around rank k) again with kc = 112;
// (3) what still fails through the brute-force scan.  Every tier returns the exact answer or hands the query on, so the
// result does not depend on which tier served it.
// begin = tier 1, enqueued without any host synchronisation (the count of uncertified queries travels to a pinned host word
// behind an event); finish = wait for that word and, only if it is not zero, run tiers 2 and 3 (which do synchronise).
namespace {
struct ExactScratch {
  float *cs, *ex, *eps, *fs;
  int64_t *ci, *fi;
  int32_t *flags, *flist, *flist2, *fcnt;
  void *qsub;
};
constexpr int32_t KC_SMALL = 56, KC_BIG = 112;

int exact_scratch(ahr_index *idx, int32_t nq, int32_t k, ExactScratch &x) {
  const int32_t kc1 = k <= 50 ? KC_SMALL : KC_BIG;
  const size_t row_bytes = (size_t)idx->dim * (idx->dtype == AHR_DTYPE_F32 ? 4 : 2);
  // cand scores/ids [nq,kc], exact scores [nq,kc], eps [nq], flags [nq], list [nq], count, final (scores, ids) [nq,k] of a
  // sub-batch, the gathered queries of tier 2, its failure list
  const size_t nc = (size_t)nq * KC_BIG;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
  const size_t o_cs = carve(nc * 4), o_ci = carve(nc * 8), o_ex = carve(nc * 4), o_eps = carve((size_t)nq * 4),
               o_flag = carve((size_t)nq * 4), o_list = carve((size_t)nq * 4), o_cnt = carve(4),
               o_fs = carve((size_t)nq * k * 4), o_fi = carve((size_t)nq * k * 8), o_list2 = carve((size_t)nq * 4),
               o_qsub = carve(kc1 == KC_SMALL ? (size_t)nq * row_bytes : 0);
  int rc = idx->ws2.reserve(off);
  if (rc != AHR_OK) return rc;
  uint8_t *w = reinterpret_cast<uint8_t *>(idx->ws2.ptr);
  x.cs = reinterpret_cast<float *>(w + o_cs);
  x.ci = reinterpret_cast<int64_t *>(w + o_ci);
  x.ex = reinterpret_cast<float *>(w + o_ex);
  x.eps = reinterpret_cast<float *>(w + o_eps);
  x.flags = reinterpret_cast<int32_t *>(w + o_flag);
  x.flist = reinterpret_cast<int32_t *>(w + o_list);
  x.flist2 = reinterpret_cast<int32_t *>(w + o_list2);
  x.fcnt = reinterpret_cast<int32_t *>(w + o_cnt);
  x.fs = reinterpret_cast<float *>(w + o_fs);
  x.fi = reinterpret_cast<int64_t *>(w + o_fi);
  x.qsub = w + o_qsub;
  return AHR_OK;
}

float exact_eps_rel(const ahr_index *idx) {
  // |tensor - canonical| <= eps_rel * ||q|| * ||x||.  Worst-case model of the tcgen05 accumulation (it truncates rather
  // than rounds, and aligns the 16 exact products of a K = 16 step to the largest exponent before adding them):
  //   per K-step  <= 2 ulp(acc) for the truncated add + 16 * 2^-24 * max|product| for the alignment
  //   dim/16 steps, |acc| and sum|products| <= ||q|| ||x||  ->  (dim/16) * 2 * 2^-23 + dim * 2^-24 = (dim/4) * 2^-23
  // All-positive / strongly correlated rows make the truncation errors add up with the same sign (they cancel on the
  // zero-mean rows of the synthetic corpora: measured error there is ~ bound/8), so the bound carries a further 2x:
  // (dim/2 + 64) * 2^-23.  tests/test_gpu_topk.py asserts |tc - canonical| <= eps/4 on all-positive, large-norm and
  // dim = 1024..4096 inputs.  The only cost of a wide bound is more tier-2 / brute-force fallbacks.
  // fp32 rows go through the hi/lo split: 3x the accumulation steps plus the dropped lo*lo and residual terms
  // (|x - hi - lo| <= 2^-18 |x| per operand -> <= 3 * 2^-18, rounded up to 2^-16)
  return idx->dtype == AHR_DTYPE_F32 ? (float)((3 * idx->dim / 2 + 64) * std::ldexp(1.0, -23) + std::ldexp(1.0, -16))
                                     : (float)((idx->dim / 2 + 64) * std::ldexp(1.0, -23));
}

// one tier over `n` queries at `q`, enqueued only: results into out_* [n, k], uncertified query indices into `fail_list`,
// their number into x.fcnt
int exact_tier(ahr_index *idx, const ExactScratch &x, const void *q, int32_t n, int32_t kc, int32_t k, float *out_s,
               int64_t *out_i, int32_t *fail_list, cudaStream_t st) {
  int r = tc_topk(idx, q, n, kc, x.cs, x.ci, nullptr, st);
  if (r == AHR_OK) r = exact_rescore(idx->rows, q, idx->dtype, idx->dim, n, kc, x.ci, idx->id_base, x.ex, st);
  if (r == AHR_OK) r = query_eps(q, idx->dtype, n, idx->dim, idx->d_max_norm, exact_eps_rel(idx), x.eps, st);
  if (r != AHR_OK) return r;
  AHR_CUDA_CHECK(cudaMemsetAsync(x.fcnt, 0, 4, st));
  return exact_certify(x.cs, x.ci, x.ex, n, kc, k, x.eps, out_s, out_i, x.flags, x.fcnt, fail_list, st);
}
}  // namespace

int ahr_topk_exact_begin(ahr_index *idx, const void *dev_q, int32_t nq, int32_t k, float *dev_scores, int64_t *dev_ids,
                         void *stream) {
  AHR_NVTX("ahr K1x exact mode: tensor-core candidates, rescoring, certification");
  AHR_REQUIRE(idx != nullptr, "ahr_topk_exact_begin: index is NULL");
  AHR_REQUIRE(!idx->pending.active, "ahr_topk_exact_begin: the previous exact query of this index was not finished");
  AHR_REQUIRE(nq >= 0 && k >= 1 && k <= SELECT_MAX_K, "ahr_topk_exact_begin: bad sizes nq=%d k=%d", nq, k);
  if (nq == 0) return AHR_OK;
  AHR_REQUIRE(dev_q && dev_scores && dev_ids, "ahr_topk_exact_begin: NULL buffer");
  DeviceGuard guard(idx->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  idx->n_certified = 0;
  idx->n_fallback = 0;
  auto &pd = idx->pending;
  pd = ahr_index::Pending();
  pd.q = dev_q; pd.nq = nq; pd.k = k; pd.scores = dev_scores; pd.ids = dev_ids; pd.stream = st;
  pd.use_tc = idx->has_map && k <= 96 && idx->n_rows > 4096 && (reinterpret_cast<uintptr_t>(dev_q) & 15) == 0;
  pd.active = true;
  if (!pd.use_tc) {  // shapes the tensor path does not take: the brute-force scan, enqueued here, nothing left for finish
    idx->n_fallback = nq;
    int rc = exact_bruteforce(idx, dev_q, nullptr, nq, k, dev_scores, dev_ids, idx->ws, st);
    pd.active = rc == AHR_OK;
    pd.done = true;
    return rc;
  }
  ExactScratch x;
  int rc = exact_scratch(idx, nq, k, x);
  if (rc == AHR_OK) rc = exact_tier(idx, x, dev_q, nq, k <= 50 ? KC_SMALL : KC_BIG, k, dev_scores, dev_ids, x.flist, st);
  if (rc != AHR_OK) { pd.active = false; return rc; }
  *idx->h_nfail = -1;
  AHR_CUDA_CHECK(cudaMemcpyAsync(idx->h_nfail, x.fcnt, 4, cudaMemcpyDeviceToHost, st));  // pinned destination
  AHR_CUDA_CHECK(cudaEventRecord(idx->ev_nfail, st));
  return AHR_OK;
}

int ahr_topk_exact_finish(ahr_index *idx) {
  AHR_REQUIRE(idx != nullptr, "ahr_topk_exact_finish: index is NULL");
  auto &pd = idx->pending;
  if (!pd.active) return AHR_OK;  // nothing pending (nq == 0)
  pd.active = false;
  if (pd.done) return AHR_OK;
  DeviceGuard guard(idx->device);
  cudaStream_t st = pd.stream;
  const int32_t nq = pd.nq, k = pd.k;
  AHR_CUDA_CHECK(cudaEventSynchronize(idx->ev_nfail));  // the one wait of the common case: tier 1 has run
  int32_t n_fail = *idx->h_nfail;
  if (n_fail == 0) {
    idx->n_certified = nq;
    return AHR_OK;
  }
  AHR_NVTX("ahr K1x exact mode: fallback tiers");
  ExactScratch x;
  int rc = exact_scratch(idx, nq, k, x);  // same layout as in begin (the workspace does not move: same size)
  if (rc != AHR_OK) return rc;
  const size_t row_bytes = (size_t)idx->dim * (idx->dtype == AHR_DTYPE_F32 ? 4 : 2);
  if (k <= 50) {
    // tier 2 on the sub-batch of uncertified queries
    const int32_t n1 = n_fail;
    rc = gather_query_rows(pd.q, (int64_t)row_bytes, x.flist, n1, x.qsub, st);
    if (rc == AHR_OK) rc = exact_tier(idx, x, x.qsub, n1, KC_BIG, k, x.fs, x.fi, x.flist2, st);
    if (rc != AHR_OK) return rc;
    int32_t n2 = 0;
    AHR_CUDA_CHECK(cudaMemcpyAsync(&n2, x.fcnt, 4, cudaMemcpyDeviceToHost, st));
    AHR_CUDA_CHECK(cudaStreamSynchronize(st));
    // rows of tier-2 failures are overwritten by tier 3 below
    rc = gather_rows_i64(x.fs, x.fi, x.flist, n1, k, pd.scores, pd.ids, /*scatter=*/true, st);
    if (rc != AHR_OK) return rc;
    if (n2 > 0) {  // compose the failure list: positions in the sub-batch -> original query indices
      std::vector<int32_t> h1(n1), h2(n2);
      AHR_CUDA_CHECK(cudaMemcpyAsync(h1.data(), x.flist, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
      AHR_CUDA_CHECK(cudaMemcpyAsync(h2.data(), x.flist2, (size_t)n2 * 4, cudaMemcpyDeviceToHost, st));
      AHR_CUDA_CHECK(cudaStreamSynchronize(st));
      for (int32_t i = 0; i < n2; ++i) h2[i] = h1[h2[i]];
      AHR_CUDA_CHECK(cudaMemcpyAsync(x.flist, h2.data(), (size_t)n2 * 4, cudaMemcpyHostToDevice, st));
      AHR_CUDA_CHECK(cudaStreamSynchronize(st));  // h2 goes out of scope
    }
    n_fail = n2;
  }
  idx->n_certified = nq - n_fail;
  idx->n_fallback = n_fail;
  if (n_fail > 0) {
    rc = exact_bruteforce(idx, pd.q, x.flist, n_fail, k, x.fs, x.fi, idx->ws, st);
    if (rc != AHR_OK) return rc;
    rc = gather_rows_i64(x.fs, x.fi, x.flist, n_fail, k, pd.scores, pd.ids, /*scatter=*/true, st);
    if (rc != AHR_OK) return rc;
  }
  return AHR_OK;
}

int ahr_topk_host(ahr_index *idx, const void *host_q, int32_t nq, int32_t k, int mode, float *host_scores,
                  int64_t *host_ids, void *stream) {
  AHR_REQUIRE(idx != nullptr, "ahr_topk_host: index is NULL");
  AHR_REQUIRE(nq >= 0 && k >= 1, "ahr_topk_host: bad sizes nq=%d k=%d", nq, k);
  if (nq == 0) return AHR_OK;
  AHR_REQUIRE(host_q && host_scores && host_ids, "ahr_topk_host: NULL buffer");
  DeviceGuard guard(idx->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t esz = idx->dtype == AHR_DTYPE_F32 ? 4 : 2;
  const size_t qb = (size_t)nq * idx->dim * esz, sb = (size_t)nq * k * 4, ib = (size_t)nq * k * 8;
  int rc;
  if ((rc = idx->hq.reserve(qb)) != AHR_OK) return rc;
  if ((rc = idx->hs.reserve(sb)) != AHR_OK) return rc;
  if ((rc = idx->hi.reserve(ib)) != AHR_OK) return rc;
  AHR_CUDA_CHECK(cudaMemcpyAsync(idx->hq.ptr, host_q, qb, cudaMemcpyHostToDevice, st));
  rc = ahr_topk(idx, idx->hq.ptr, nq, k, mode, reinterpret_cast<float *>(idx->hs.ptr),
                reinterpret_cast<int64_t *>(idx->hi.ptr), stream);
  if (rc != AHR_OK) return rc;
  AHR_CUDA_CHECK(cudaMemcpyAsync(host_scores, idx->hs.ptr, sb, cudaMemcpyDeviceToHost, st));
  AHR_CUDA_CHECK(cudaMemcpyAsync(host_ids, idx->hi.ptr, ib, cudaMemcpyDeviceToHost, st));
  AHR_CUDA_CHECK(cudaStreamSynchronize(st));
  return AHR_OK;
}

int ahr_merge_topk(int device, const float *dev_scores, const int64_t *dev_ids, int32_t parts, int32_t nq, int32_t k,
                   float *dev_out_scores, int64_t *dev_out_ids, void *stream) {
  AHR_REQUIRE(parts >= 1 && nq >= 0 && k >= 1, "ahr_merge_topk: bad sizes parts=%d nq=%d k=%d", parts, nq, k);
  if (nq == 0) return AHR_OK;
  AHR_REQUIRE(dev_scores && dev_ids && dev_out_scores && dev_out_ids, "ahr_merge_topk: NULL buffer");
  DeviceGuard guard(device);
  // [parts, nq, k]: candidate i = (part, j) of query q at q * k + part * (nq * k) + j.  The lists are the sorted outputs
  // of ahr_topk (documented contract), so a few of them merge by rank counting.
  if (merge_sorted_applicable(parts, k, k))
    return merge_sorted_lists(dev_scores, nullptr, dev_ids, k, parts, k, (int64_t)nq * k, (int64_t)nq * k, nq, k, 0,
                              dev_out_scores, dev_out_ids, nullptr, reinterpret_cast<cudaStream_t>(stream));
  return select_topk_strided(dev_scores, nullptr, dev_ids, k, (int64_t)parts * k, k, (int64_t)nq * k, nq, k, 0, false,
                             dev_out_scores, dev_out_ids, reinterpret_cast<cudaStream_t>(stream));
}

int ahr_merge_topk_packed(int device, const void *dev_packed, int64_t record_bytes, int32_t parts, int32_t nq, int32_t k,
                          float *dev_out_scores, int64_t *dev_out_ids, void *stream) {
  AHR_REQUIRE(parts >= 1 && nq >= 0 && k >= 1, "ahr_merge_topk_packed: bad sizes parts=%d nq=%d k=%d", parts, nq, k);
  if (nq == 0) return AHR_OK;
  const int64_t sb = (int64_t)nq * k * 4;
  AHR_REQUIRE(dev_packed && dev_out_scores && dev_out_ids, "ahr_merge_topk_packed: NULL buffer");
  AHR_REQUIRE(sb % 8 == 0 && record_bytes % 8 == 0 && record_bytes >= sb * 3,
              "ahr_merge_topk_packed: record of %lld bytes cannot hold %d x %d (score, id) pairs 8-byte aligned",
              (long long)record_bytes, nq, k);
  DeviceGuard guard(device);
  const float *scores = reinterpret_cast<const float *>(dev_packed);
  const int64_t *ids = reinterpret_cast<const int64_t *>(reinterpret_cast<const uint8_t *>(dev_packed) + sb);
  // part p: scores at + p * record_bytes / 4 floats, ids at + p * record_bytes / 8 int64
  if (merge_sorted_applicable(parts, k, k))
    return merge_sorted_lists(scores, nullptr, ids, k, parts, k, record_bytes / 4, record_bytes / 8, nq, k, 0, dev_out_scores,
                              dev_out_ids, nullptr, reinterpret_cast<cudaStream_t>(stream));
  return select_topk_strided2(scores, nullptr, ids, k, (int64_t)parts * k, k, record_bytes / 4, record_bytes / 8, nq, k, 0,
                              false, dev_out_scores, dev_out_ids, reinterpret_cast<cudaStream_t>(stream));
}

int ahr_debug_tc_scores(ahr_index *idx, const void *dev_q, int32_t nq, float *dev_out, void *stream) {
  AHR_REQUIRE(idx && dev_q && dev_out && nq >= 1, "ahr_debug_tc_scores: bad arguments");
  AHR_REQUIRE(idx->n_rows >= 1, "ahr_debug_tc_scores: empty index");
  DeviceGuard guard(idx->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int k = 8;
  const size_t ids_off = (((size_t)nq * k * 4) + 255) & ~(size_t)255;
  int rc = idx->ws2.reserve(ids_off + (size_t)nq * k * 8);
  if (rc != AHR_OK) return rc;
  float *s = reinterpret_cast<float *>(idx->ws2.ptr);
  int64_t *i = reinterpret_cast<int64_t *>(reinterpret_cast<uint8_t *>(idx->ws2.ptr) + ids_off);
  return tc_topk(idx, dev_q, nq, k, s, i, dev_out, st);
}

int ahr_rescore(ahr_index *idx, const void *dev_q, int32_t nq, int32_t m, const int64_t *dev_cand_ids,
                float *dev_out_scores, void *stream) {
  AHR_REQUIRE(idx && dev_q && dev_cand_ids && dev_out_scores, "ahr_rescore: NULL argument");
  DeviceGuard guard(idx->device);
  return exact_rescore(idx->rows, dev_q, idx->dtype, idx->dim, nq, m, dev_cand_ids, idx->id_base, dev_out_scores,
                       reinterpret_cast<cudaStream_t>(stream));
}

int ahr_exact_stats(ahr_index *idx, int64_t *n_certified, int64_t *n_fallback) {
  AHR_REQUIRE(idx != nullptr, "ahr_exact_stats: index is NULL");
  if (n_certified) *n_certified = idx->n_certified;
  if (n_fallback) *n_fallback = idx->n_fallback;
  return AHR_OK;
}

int ahr_set_profiling(ahr_index *idx, int enable) {
  AHR_REQUIRE(idx != nullptr, "ahr_set_profiling: index is NULL");
  idx->profile = enable != 0;
  return AHR_OK;
}

int ahr_get_profile(ahr_index *idx, double *k1_ms_total, int64_t *k1_launches, int32_t *n_slabs) {
  AHR_REQUIRE(idx != nullptr, "ahr_get_profile: index is NULL");
  DeviceGuard guard(idx->device);
  double total = 0.0;
  for (auto &pr : idx->prof_events) {
    AHR_CUDA_CHECK(cudaEventSynchronize(pr.second));
    float ms = 0.f;
    AHR_CUDA_CHECK(cudaEventElapsedTime(&ms, pr.first, pr.second));
    total += ms;
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  if (k1_ms_total) *k1_ms_total = total;
  if (k1_launches) *k1_launches = (int64_t)idx->prof_events.size();
  if (n_slabs) *n_slabs = idx->last_slabs;
  idx->prof_events.clear();
  return AHR_OK;
}

int ahr_debug_round4(int device, const double *dev_x, int64_t n, double *dev_out, void *stream) {
  AHR_REQUIRE(dev_x && dev_out, "ahr_debug_round4: NULL buffer");
  DeviceGuard guard(device);
  return debug_round4(dev_x, n, dev_out, reinterpret_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------
int ahr_graph_create(ahr_graph **out, int device, int64_t n_nodes, const int64_t *const dev_row_ptr[AHR_N_CSR],
                     const int32_t *const dev_col_idx[AHR_N_CSR], const uint8_t *dev_node_type,
                     const uint8_t *dev_level, const double *dev_judge, const double *dev_judge_term,
                     const double *dev_conf) {
  AHR_REQUIRE(out != nullptr, "ahr_graph_create: out is NULL");
  AHR_REQUIRE(n_nodes >= 0 && n_nodes < (int64_t)INT32_MAX, "ahr_graph_create: n_nodes %lld out of range", (long long)n_nodes);
  AHR_REQUIRE(dev_row_ptr && dev_col_idx, "ahr_graph_create: NULL CSR tables");
  for (int s = 0; s < AHR_N_CSR; ++s) AHR_REQUIRE(dev_row_ptr[s] != nullptr, "ahr_graph_create: row_ptr[%d] is NULL", s);
  AHR_REQUIRE(n_nodes == 0 || (dev_node_type && dev_level && dev_judge && dev_judge_term && dev_conf),
              "ahr_graph_create: NULL attribute array");
  DeviceGuard guard(device);
  if (!guard.ok) {
    set_error("ahr_graph_create: cannot select CUDA device %d", device);
    return AHR_E_CUDA;
  }
  ahr_graph *g = new ahr_graph();
  g->device = device;
  g->view.n_nodes = n_nodes;
  for (int s = 0; s < AHR_N_CSR; ++s) {
    g->view.row_ptr[s] = dev_row_ptr[s];
    g->view.col_idx[s] = dev_col_idx[s];
  }
  g->view.line = nullptr;
  if (n_nodes > 0) {
    // pack the five attribute arrays + the head of each node's anchor list into one 64-byte line per node (the caller's
    // attribute arrays are not referenced afterwards)
    int rc = g->recs.reserve((size_t)n_nodes * sizeof(NodeLine));
    if (rc == AHR_OK)
      rc = graph_pack_nodes(n_nodes, dev_node_type, dev_level, dev_judge, dev_judge_term, dev_conf, dev_row_ptr, dev_col_idx,
                            reinterpret_cast<NodeLine *>(g->recs.ptr), 0);
    if (rc == AHR_OK && cudaStreamSynchronize(0) != cudaSuccess) {
      set_error("ahr_graph_create: packing the node attributes failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = AHR_E_CUDA;
    }
    if (rc != AHR_OK) { g->recs.release(); delete g; return rc; }
    g->view.line = reinterpret_cast<const NodeLine *>(g->recs.ptr);
  }
  *out = g;
  return AHR_OK;
}

int ahr_graph_destroy(ahr_graph *g) {
  if (!g) return AHR_OK;
  DeviceGuard guard(g->device);
  g->ws.release();
  g->recs.release();
  delete g;
  return AHR_OK;
}

static int anchor_rerank_common(ahr_graph *g, int32_t nq, int32_t n_seeds, const int32_t *dev_seed_nodes,
                                const int64_t *dev_seed_rows, const int32_t *dev_row_to_node, int64_t n_rows,
                                const float *dev_seed_scores, const ahr_rerank_params *params, int32_t top_k,
                                int32_t *dev_out_nodes, double *dev_out_semantic, double *dev_out_score,
                                int32_t *dev_out_count, int32_t *dev_exp_nodes, double *dev_exp_semantic,
                                uint8_t *dev_exp_kind, int32_t *dev_exp_count, double *dev_seed_semantic, void *stream) {
  AHR_NVTX("ahr K2 anchor expansion + rerank");
  AHR_REQUIRE(g && params, "ahr_anchor_rerank: NULL handle or params");
  if (nq == 0) return AHR_OK;
  AHR_REQUIRE(dev_seed_scores && dev_out_nodes && dev_out_semantic && dev_out_score && dev_out_count,
              "ahr_anchor_rerank: NULL buffer");
  AHR_REQUIRE(!dev_exp_nodes || dev_exp_semantic, "ahr_anchor_rerank: exp_nodes needs exp_semantic");
  DeviceGuard guard(g->device);
  const size_t need = graph_anchor_workspace_bytes(nq, n_seeds, *params);
  if (need) {
    int rc = g->ws.reserve(need);
    if (rc != AHR_OK) return rc;
  }
  return graph_anchor_rerank(g->view, nq, n_seeds, dev_seed_nodes, dev_seed_rows, dev_row_to_node, n_rows, dev_seed_scores,
                             *params, top_k, dev_out_nodes, dev_out_semantic, dev_out_score, dev_out_count, dev_exp_nodes,
                             dev_exp_semantic, dev_exp_kind, dev_exp_count, dev_seed_semantic, g->ws.ptr, g->ws.bytes,
                             reinterpret_cast<cudaStream_t>(stream));
}

int ahr_anchor_rerank(ahr_graph *g, int32_t nq, int32_t n_seeds, const int32_t *dev_seed_nodes,
                      const float *dev_seed_scores, const ahr_rerank_params *params, int32_t top_k,
                      int32_t *dev_out_nodes, double *dev_out_semantic, double *dev_out_score, int32_t *dev_out_count,
                      int32_t *dev_exp_nodes, double *dev_exp_semantic, uint8_t *dev_exp_kind, int32_t *dev_exp_count,
                      double *dev_seed_semantic, void *stream) {
  AHR_REQUIRE(nq == 0 || dev_seed_nodes, "ahr_anchor_rerank: NULL seed list");
  return anchor_rerank_common(g, nq, n_seeds, dev_seed_nodes, nullptr, nullptr, 0, dev_seed_scores, params, top_k,
                              dev_out_nodes, dev_out_semantic, dev_out_score, dev_out_count, dev_exp_nodes, dev_exp_semantic,
                              dev_exp_kind, dev_exp_count, dev_seed_semantic, stream);
}

int ahr_anchor_rerank_rows(ahr_graph *g, int32_t nq, int32_t n_seeds, const int64_t *dev_seed_rows,
                           const int32_t *dev_row_to_node, int64_t n_rows, const float *dev_seed_scores,
                           const ahr_rerank_params *params, int32_t top_k, int32_t *dev_out_nodes,
                           double *dev_out_semantic, double *dev_out_score, int32_t *dev_out_count, int32_t *dev_exp_nodes,
                           double *dev_exp_semantic, uint8_t *dev_exp_kind, int32_t *dev_exp_count,
                           double *dev_seed_semantic, void *stream) {
  AHR_REQUIRE(nq == 0 || (dev_seed_rows && dev_row_to_node && n_rows >= 0), "ahr_anchor_rerank_rows: NULL seed rows / table");
  return anchor_rerank_common(g, nq, n_seeds, nullptr, dev_seed_rows, dev_row_to_node, n_rows, dev_seed_scores, params, top_k,
                              dev_out_nodes, dev_out_semantic, dev_out_score, dev_out_count, dev_exp_nodes, dev_exp_semantic,
                              dev_exp_kind, dev_exp_count, dev_seed_semantic, stream);
}

int ahr_expand(ahr_graph *g, int32_t n_env, const int32_t *dev_ops, const int32_t *dev_in_ids,
               const int32_t *dev_in_offsets, int32_t limit, int32_t *dev_out_ids, int32_t *dev_out_counts,
               void *stream) {
  AHR_NVTX("ahr K2 frontier expansion");
  AHR_REQUIRE(g != nullptr, "ahr_expand: graph is NULL");
  if (n_env == 0) return AHR_OK;
  AHR_REQUIRE(dev_ops && dev_in_offsets && dev_out_ids && dev_out_counts, "ahr_expand: NULL buffer");
  DeviceGuard guard(g->device);
  return graph_expand(g->view, n_env, dev_ops, dev_in_ids, dev_in_offsets, limit, dev_out_ids, dev_out_counts,
                      reinterpret_cast<cudaStream_t>(stream));
}

int ahr_lca_closure(ahr_graph *g, const int32_t *host_ids, int32_t n_ids, uint8_t *dev_out_flags, int64_t *n_result,
                    void *stream) {
  AHR_REQUIRE(g && dev_out_flags, "ahr_lca_closure: NULL argument");
  AHR_REQUIRE(n_ids >= 0 && (n_ids == 0 || host_ids != nullptr), "ahr_lca_closure: bad input list");
  for (int32_t i = 0; i < n_ids; ++i)
    AHR_REQUIRE(host_ids[i] >= 0 && host_ids[i] < g->view.n_nodes, "ahr_lca_closure: node id %d out of range", host_ids[i]);
  DeviceGuard guard(g->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (n_result) *n_result = 0;
  if (g->view.n_nodes == 0) return AHR_OK;
  if (n_ids == 0) {  // set.intersection() of nothing: the reference returns no node (environment.py:276)
    AHR_CUDA_CHECK(cudaMemsetAsync(dev_out_flags, 0, (size_t)g->view.n_nodes, st));
    return AHR_OK;
  }
  int rc = g->ws.reserve(graph_lca_workspace_bytes(g->view.n_nodes, n_ids));
  if (rc != AHR_OK) return rc;
  return graph_lca_closure(g->view, host_ids, n_ids, dev_out_flags, n_result, g->ws.ptr, st);
}

int ahr_env_reset(ahr_graph *g, const ahr_env_batch *env, const int32_t *dev_nodes, const double *dev_semantic,
                  const double *dev_score, const int32_t *dev_count, int32_t top_k, int32_t n_seeds, float *dev_features,
                  uint8_t *dev_mask, void *stream) {
  AHR_NVTX("ahr gym reset");
  AHR_REQUIRE(g && env, "ahr_env_reset: NULL handle or env");
  AHR_REQUIRE(top_k >= 1 && top_k <= 10 && n_seeds >= 0, "ahr_env_reset: top_k must be in [1, 10] (got %d)", top_k);
  if (env->n_env == 0) return AHR_OK;
  AHR_REQUIRE(dev_nodes && dev_semantic && dev_score && dev_count && dev_features && dev_mask, "ahr_env_reset: NULL buffer");
  DeviceGuard guard(g->device);
  return env_reset(g->view, *env, dev_nodes, dev_semantic, dev_score, dev_count, top_k, n_seeds, dev_features, dev_mask,
                   reinterpret_cast<cudaStream_t>(stream));
}

int ahr_env_step(ahr_graph *g, const ahr_env_batch *env, const int32_t *dev_actions, int32_t max_steps,
                 double repeat_penalty, float *dev_features, double *dev_rewards, uint8_t *dev_done, uint8_t *dev_mask,
                 void *stream) {
  AHR_NVTX("ahr gym step");
  AHR_REQUIRE(g && env, "ahr_env_step: NULL handle or env");
  if (env->n_env == 0) return AHR_OK;
  AHR_REQUIRE(dev_actions && dev_features && dev_rewards && dev_done && dev_mask, "ahr_env_step: NULL buffer");
  DeviceGuard guard(g->device);
  return env_step(g->view, *env, dev_actions, max_steps, repeat_penalty, dev_features, dev_rewards, dev_done, dev_mask,
                  reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
