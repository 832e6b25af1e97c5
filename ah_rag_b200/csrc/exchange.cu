// exchange.cu -- K3: the cross-GPU candidate exchange of the sharded index, over NVLink peer memory.
//
// The reference has no multi-GPU path; the shape of the exchange comes from the north star:
// every rank computes a local top-k over its node range, the [nq, k] candidate lists of all ranks meet, a device
// k-way merge with the total order (score desc, id asc) produces the same answer on every rank.
//
// Instead of a collective library call the lists travel by plain stores into the PEERS' memory:
//
//   buffer (one cudaMalloc per rank, exported with cudaIpcGetMemHandle, opened by every peer):
//     flags [2][world]  uint64   epoch of the last record rank `src` delivered into half `h`
//     slots [2][world][record]   record = [nq*k fp32 scores | nq*k int64 ids], slot `src` is written by rank src
//
//   push_kernel   copies this rank's record (already sitting in its own slot: ahr_topk wrote it there) into slot
//                 `rank` of every peer with 16-byte stores over NVLink, fences at system scope, and the LAST block to
//                 finish publishes the epoch into the peers' flags (st.release.sys).
//   wait + merge  one warp acquires the world flags (ld.acquire.sys spin), then select_topk_kernel merges the local
//                 slots in stream order.  No host synchronisation, no collective library call.
//
// Two halves alternate by epoch parity: a rank can run at most one exchange ahead of a peer (its merge of call n
// needs the peer's record of call n), so the half a fast rank overwrites in call n+1 is one the slow rank finished
// reading in call n-1.
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace ahr {

constexpr int PUSH_THREADS = 256;
constexpr int PUSH_BLOCKS_PER_PEER = 20;  // 140 blocks at 8 GPUs: about one per SM
constexpr int MAX_WORLD = 16;

__device__ __forceinline__ void st_release_sys_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t *p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(PUSH_THREADS)
push_kernel(PeerPush peers, size_t rec_bytes) {
  const int rank = peers.rank, world = peers.world;
  const uint8_t *src = peers.base[rank] + peers.slot_off;
  const size_t n16 = rec_bytes / 16;
  // every peer gets its own group of blocks, so all the peers are written concurrently (a loop over the peers left the
  // transfer latency-bound on one destination at a time: 25 us per peer, 175 us at 8 GPUs)
  const uint4 *src4 = reinterpret_cast<const uint4 *>(src);
  const int n_peers = world - 1;
  const int per_peer = gridDim.x / n_peers;            // blocks per peer (grid is a multiple of n_peers)
  const int pi = blockIdx.x / per_peer, sub = blockIdx.x - pi * per_peer;
  if (pi < n_peers) {
    const int peer = pi < rank ? pi : pi + 1;
    uint4 *dst4 = reinterpret_cast<uint4 *>(peers.base[peer] + peers.slot_off);  // the peer's slot `rank`, over NVLink
    const size_t stride = (size_t)per_peer * PUSH_THREADS;
    size_t i = (size_t)sub * PUSH_THREADS + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {  // four independent loads in flight per thread
      const uint4 v0 = src4[i], v1 = src4[i + stride], v2 = src4[i + 2 * stride], v3 = src4[i + 3 * stride];
      dst4[i] = v0; dst4[i + stride] = v1; dst4[i + 2 * stride] = v2; dst4[i + 3 * stride] = v3;
    }
    for (; i < n16; i += stride) dst4[i] = src4[i];
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(peers.done_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence_system();
  if (threadIdx.x < world && (int)threadIdx.x != rank)
    st_release_sys_u64(reinterpret_cast<uint64_t *>(peers.base[threadIdx.x] + peers.flag_off) + rank, peers.epoch);
  if (threadIdx.x == 0) *peers.done_counter = 0;  // ready for the next launch (stream-ordered)
}

// a rank that failed after it entered an exchange tells its peers so: they stop at once instead of waiting out the timeout
constexpr uint64_t EPOCH_POISON = ~0ull;
__global__ void poison_kernel(PeerPush peers) {
  if ((int)threadIdx.x < peers.world && (int)threadIdx.x != peers.rank)
    st_release_sys_u64(reinterpret_cast<uint64_t *>(peers.base[threadIdx.x] + peers.flag_off) + peers.rank, EPOCH_POISON);
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Blocks until every peer's record of this epoch has landed (one acquire per peer, by the first `world` threads).
// A peer that never delivers (crashed rank) must not leave this GPU spinning for ever, nor let the merge run on stale
// data: after `timeout_ns` the kernel traps, which surfaces as a CUDA error on the next call of this process.
__global__ void wait_flags_kernel(const uint64_t *flags, int rank, int world, uint64_t epoch, unsigned long long timeout_ns) {
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    const unsigned long long t0 = global_timer_ns();
    uint64_t seen;
    while ((seen = ld_acquire_sys_u64(flags + threadIdx.x)) < epoch) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > timeout_ns) __trap();
    }
    if (seen == EPOCH_POISON) __trap();  // the peer failed inside this exchange: fail now, never merge stale data
  }
}

}  // namespace ahr

using namespace ahr;

struct ahr_comm {
  int device = 0, rank = 0, world = 1;
  size_t max_record = 0, flags_bytes = 0, total_bytes = 0;
  uint8_t *base = nullptr;               // own allocation
  std::vector<uint8_t *> peer;           // opened peer allocations (peer[rank] = base)
  std::vector<bool> opened;
  unsigned int *done_counter = nullptr;
  uint64_t epoch = 0;
  bool connected = false;
  unsigned long long timeout_ns = 600ull * 1000000000ull;  // AHR_COMM_TIMEOUT_S (default 600 s, like torch's NCCL watchdog)
  size_t slot_offset(int half, int src) const { return flags_bytes + ((size_t)half * world + src) * max_record; }
  size_t flag_offset(int half) const { return (size_t)half * world * sizeof(uint64_t); }
};

extern "C" {

int ahr_comm_create(ahr_comm **out, int device, int rank, int world, int64_t max_record_bytes) {
  AHR_REQUIRE(out != nullptr, "ahr_comm_create: out is NULL");
  AHR_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "ahr_comm_create: bad rank %d / world %d", rank, world);
  AHR_REQUIRE(max_record_bytes >= 16, "ahr_comm_create: record too small");
  DeviceGuard guard(device);
  AHR_REQUIRE(guard.ok, "ahr_comm_create: cannot select CUDA device %d", device);
  ahr_comm *c = new ahr_comm();
  c->device = device; c->rank = rank; c->world = world;
  c->max_record = ((size_t)max_record_bytes + 255) & ~(size_t)255;
  c->flags_bytes = ((size_t)2 * world * sizeof(uint64_t) + 255) & ~(size_t)255;
  c->total_bytes = c->flags_bytes + (size_t)2 * world * c->max_record;
  cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&c->base), c->total_bytes);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&c->done_counter), sizeof(unsigned int));
  if (e != cudaSuccess) {
    set_error("ahr_comm_create: cudaMalloc(%zu) failed: %s", c->total_bytes, cudaGetErrorString(e));
    if (c->base) cudaFree(c->base);
    delete c;
    return AHR_E_NOMEM;
  }
  AHR_CUDA_CHECK(cudaMemset(c->base, 0, c->flags_bytes));
  AHR_CUDA_CHECK(cudaMemset(c->done_counter, 0, sizeof(unsigned int)));
  AHR_CUDA_CHECK(cudaDeviceSynchronize());
  c->peer.assign(world, nullptr);
  c->opened.assign(world, false);
  c->peer[rank] = c->base;
  if (const char *e = getenv("AHR_COMM_TIMEOUT_S"))
    if (atof(e) > 0) c->timeout_ns = (unsigned long long)(atof(e) * 1e9);
  c->connected = world == 1;
  *out = c;
  return AHR_OK;
}

int ahr_comm_export(ahr_comm *c, void *handle_out) {
  AHR_REQUIRE(c && handle_out, "ahr_comm_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == AHR_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard guard(c->device);
  cudaIpcMemHandle_t h;
  AHR_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->base));
  memcpy(handle_out, &h, sizeof h);
  return AHR_OK;
}

int ahr_comm_connect(ahr_comm *c, const void *all_handles) {
  AHR_REQUIRE(c && all_handles, "ahr_comm_connect: NULL argument");
  DeviceGuard guard(c->device);
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank || c->opened[p]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const uint8_t *>(all_handles) + (size_t)p * AHR_IPC_HANDLE_BYTES, sizeof h);
    void *ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("ahr_comm_connect: cudaIpcOpenMemHandle for rank %d failed: %s", p, cudaGetErrorString(e));
      cudaGetLastError();
      return AHR_E_CUDA;
    }
    c->peer[p] = reinterpret_cast<uint8_t *>(ptr);
    c->opened[p] = true;
  }
  c->connected = true;
  return AHR_OK;
}

int ahr_comm_destroy(ahr_comm *c) {
  if (!c) return AHR_OK;
  DeviceGuard guard(c->device);
  cudaDeviceSynchronize();
  for (int p = 0; p < c->world; ++p)
    if (c->opened[p]) cudaIpcCloseMemHandle(c->peer[p]);
  if (c->base) cudaFree(c->base);
  if (c->done_counter) cudaFree(c->done_counter);
  delete c;
  return AHR_OK;
}

static PeerPush make_push(const ahr_comm *c, uint64_t epoch, size_t scores_bytes) {
  PeerPush p;
  for (int r = 0; r < c->world && r < 16; ++r) p.base[r] = c->peer[r];
  p.rank = c->rank;
  p.world = c->world;
  const int half = (int)(epoch & 1);
  p.slot_off = c->slot_offset(half, c->rank);
  p.flag_off = c->flag_offset(half);
  p.scores_bytes = scores_bytes;
  p.epoch = epoch;
  p.done_counter = c->done_counter;
  return p;
}

}  // extern "C"

namespace ahr {
int peer_push_record(const PeerPush &push, size_t rec_bytes, cudaStream_t stream) {
  if (push.world <= 1) return AHR_OK;
  AHR_NVTX("ahr K3 peer push");
  push_kernel<<<PUSH_BLOCKS_PER_PEER * (push.world - 1), PUSH_THREADS, 0, stream>>>(push, rec_bytes);
  AHR_KERNEL_CHECK();
  return AHR_OK;
}
}  // namespace ahr

extern "C" {

// wait for every peer's record of `epoch`, then merge the world slots of that half into out
static int wait_and_merge(ahr_comm *c, uint64_t epoch, int32_t nq, int32_t k, float *dev_out_scores, int64_t *dev_out_ids,
                          cudaStream_t st) {
  AHR_NVTX("ahr K3 wait for peers + merge");
  const size_t sb = (size_t)nq * k * 4;
  const int half = (int)(epoch & 1);
  if (c->world > 1) {
    wait_flags_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const uint64_t *>(c->base + c->flag_offset(half)), c->rank,
                                        c->world, epoch, c->timeout_ns);
    AHR_KERNEL_CHECK();
  }
  // k-way merge over the world slots of this half: part p at + p * max_record.  The slots are sorted lists.
  const float *scores = reinterpret_cast<const float *>(c->base + c->slot_offset(half, 0));
  const int64_t *ids = reinterpret_cast<const int64_t *>(c->base + c->slot_offset(half, 0) + sb);
  if (merge_sorted_applicable(c->world, k, k))
    return merge_sorted_lists(scores, nullptr, ids, k, c->world, k, (int64_t)(c->max_record / 4), (int64_t)(c->max_record / 8),
                              nq, k, 0, dev_out_scores, dev_out_ids, nullptr, st);
  return select_topk_strided2(scores, nullptr, ids, k, (int64_t)c->world * k, k, (int64_t)(c->max_record / 4),
                              (int64_t)(c->max_record / 8), nq, k, 0, false, dev_out_scores, dev_out_ids, st);
}

static int check_exchange_shape(const ahr_comm *c, int32_t nq, int32_t k, const char *who) {
  AHR_REQUIRE(c->connected, "%s: peers are not connected (ahr_comm_connect)", who);
  AHR_REQUIRE(nq >= 1 && k >= 1, "%s: bad sizes nq=%d k=%d", who, nq, k);
  const size_t sb = (size_t)nq * k * 4;
  AHR_REQUIRE(sb % 16 == 0, "%s: nq * k must be a multiple of 4 (16-byte peer stores)", who);
  AHR_REQUIRE(sb * 3 <= c->max_record, "%s: record of %zu bytes exceeds the buffer's %zu", who, sb * 3, c->max_record);
  return AHR_OK;
}

int ahr_comm_begin(ahr_comm *c, int32_t nq, int32_t k, float **dev_slot_scores, int64_t **dev_slot_ids) {
  AHR_REQUIRE(c && dev_slot_scores && dev_slot_ids, "ahr_comm_begin: NULL argument");
  int rc = check_exchange_shape(c, nq, k, "ahr_comm_begin");
  if (rc != AHR_OK) return rc;
  // the slot of the NEXT exchange; the epoch itself only advances in ahr_comm_exchange_merge, so a caller that fails
  // between the two calls has not desynchronised the ranks
  const int half = (int)((c->epoch + 1) & 1);
  uint8_t *slot = c->base + c->slot_offset(half, c->rank);
  *dev_slot_scores = reinterpret_cast<float *>(slot);
  *dev_slot_ids = reinterpret_cast<int64_t *>(slot + (size_t)nq * k * 4);
  return AHR_OK;
}

int ahr_comm_exchange_merge(ahr_comm *c, int32_t nq, int32_t k, float *dev_out_scores, int64_t *dev_out_ids, void *stream) {
  AHR_REQUIRE(c && dev_out_scores && dev_out_ids, "ahr_comm_exchange_merge: NULL argument");
  int rc = check_exchange_shape(c, nq, k, "ahr_comm_exchange_merge");
  if (rc != AHR_OK) return rc;
  DeviceGuard guard(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  c->epoch += 1;
  const PeerPush push = make_push(c, c->epoch, (size_t)nq * k * 4);
  rc = peer_push_record(push, (size_t)nq * k * 12, st);
  if (rc == AHR_OK) rc = wait_and_merge(c, c->epoch, nq, k, dev_out_scores, dev_out_ids, st);
  return rc;
}

int ahr_topk_sharded(ahr_index *idx, ahr_comm *c, const void *dev_q, int32_t nq, int32_t k, int mode, float *dev_out_scores,
                     int64_t *dev_out_ids, void *stream) {
  AHR_REQUIRE(idx && c && dev_q && dev_out_scores && dev_out_ids, "ahr_topk_sharded: NULL argument");
  AHR_REQUIRE(mode == AHR_MODE_BF16 || mode == AHR_MODE_EXACT, "ahr_topk_sharded: unknown mode %d", mode);
  int rc = check_exchange_shape(c, nq, k, "ahr_topk_sharded");  // everything that can be refused is refused BEFORE the epoch moves
  if (rc != AHR_OK) return rc;
  AHR_REQUIRE(mode != AHR_MODE_BF16 || k <= 120, "AHR_MODE_BF16 supports k <= 120 (got %d); use AHR_MODE_EXACT", k);
  DeviceGuard guard(c->device);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  c->epoch += 1;
  const uint64_t epoch = c->epoch;
  const size_t sb = (size_t)nq * k * 4;
  const PeerPush push = make_push(c, epoch, sb);
  uint8_t *slot = c->base + push.slot_off;
  float *slot_s = reinterpret_cast<float *>(slot);
  int64_t *slot_i = reinterpret_cast<int64_t *>(slot + sb);
  if (mode == AHR_MODE_BF16) {
    // the reduction of the slab lists writes this rank's slot AND the peers' copies of it, and publishes the flags
    rc = index_topk_bf16(idx, dev_q, nq, k, slot_s, slot_i, &push, st);
  } else {
    rc = ahr_topk(idx, dev_q, nq, k, mode, slot_s, slot_i, stream);
    if (rc == AHR_OK) rc = peer_push_record(push, sb * 3, st);
  }
  if (rc != AHR_OK) {
    // this rank cannot deliver: tell the peers (they trap at once instead of waiting out AHR_COMM_TIMEOUT_S)
    const std::string why = ahr_last_error();
    if (c->world > 1) poison_kernel<<<1, 32, 0, st>>>(push);
    set_error("%s", why.c_str());
    return rc;
  }
  return wait_and_merge(c, epoch, nq, k, dev_out_scores, dev_out_ids, st);
}

}  // extern "C"
