// Slab-sharded top-k: exchange + merge of the per-slab results in ONE kernel over NVLink peer memory.
//
// The index path shards the voxel rows over the GPUs of a box (one process per GPU);
// every rank ends a query batch with its slab's (nq, k) best rows.  The NCCL form of the exchange is an
// all-gather of world * nq * k * 12 bytes followed by a merge kernel.  Here the block that owns query q
//   1. stores its slab's k (global id, score) entries of q straight into every peer's receive buffer
//      (plain st.global to peer-mapped memory, CUDA IPC), fences, and raises a per-(source, query) flag on the
//      peer with a system-scope release store;
//   2. spins (system-scope acquire loads) on its OWN flags until all `world` sources have delivered q;
//   3. merges the world * k entries by (score desc, global id asc) -- the order every slab already uses.
// No host round trip, no second launch; the transfer of query q overlaps the merge of the queries whose data
// already arrived.  Receive buffers are double-buffered by epoch parity: a rank can run at most one exchange
// ahead of the slowest peer (it needs that peer's data to finish an exchange), so the slot of epoch e is never
// rewritten before every reader of epoch e - 2 is done.
//
// Replaces nothing in the reference (which is single-GPU); it serves BASELINE config 5.
#include <cstdio>
#include <cstring>

#include "avl_internal.h"

namespace avl {
namespace {

constexpr int kMaxWorld = 16;
constexpr unsigned long long kSpinTimeoutCycles = 20000000000ull;  // ~10 s: a missing peer must not hang the GPU

struct P2PView {
  uint8_t* peer[kMaxWorld];  // base of every rank's receive buffer, as mapped into THIS process
  int32_t rank, world, nq_max, k_max;
  uint64_t slot_bytes;       // one (parity, source) slot: ids then scores of nq_max * k_max entries
  uint64_t data_bytes;       // 2 * world * slot_bytes; the flags follow
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(128)
p2p_exchange_merge_kernel(const __grid_constant__ P2PView v, const int64_t* __restrict__ idx,
                          const float* __restrict__ val, int32_t nq, int32_t k, uint32_t epoch, int64_t row_offset,
                          const int64_t* __restrict__ global_ids, int64_t* __restrict__ out_idx,
                          float* __restrict__ out_val, uint32_t* __restrict__ status) {
  __shared__ int64_t si[1024];
  __shared__ float sv[1024];
  __shared__ uint32_t sh_timeout;
  if (threadIdx.x == 0) sh_timeout = 0u;
  __syncthreads();
  const int q = blockIdx.x;
  const uint32_t parity = epoch & 1u;
  const uint64_t ids_bytes = static_cast<uint64_t>(v.nq_max) * v.k_max * sizeof(int64_t);
  // ---- 1. deliver my slab's entries of query q to every rank (my own buffer included)
  for (int e = threadIdx.x; e < v.world * k; e += blockDim.x) {
    const int dst = e / k, j = e - dst * k;
    uint8_t* slot = v.peer[dst] + (static_cast<uint64_t>(parity) * v.world + v.rank) * v.slot_bytes;
    // slab-local row -> global row (empty slots stay -1): a contiguous slab adds its first row; the slab of a sharded
    // BUILD looks its rows up in the first-touch id table (ShardedBuilder.finalize)
    const int64_t id = idx[static_cast<size_t>(q) * k + j];
    reinterpret_cast<int64_t*>(slot)[static_cast<size_t>(q) * k + j] =
        id < 0 ? id : (global_ids ? global_ids[id] : id + row_offset);
    reinterpret_cast<float*>(slot + ids_bytes)[static_cast<size_t>(q) * k + j] = val[static_cast<size_t>(q) * k + j];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < v.world) {
    uint32_t* flags = reinterpret_cast<uint32_t*>(v.peer[threadIdx.x] + v.data_bytes);
    st_release_sys(flags + (static_cast<size_t>(parity) * v.world + v.rank) * v.nq_max + q, epoch);
  }
  // ---- 2. wait until every source delivered query q of this epoch
  if (threadIdx.x < v.world) {
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(v.peer[v.rank] + v.data_bytes) +
                           (static_cast<size_t>(parity) * v.world + threadIdx.x) * v.nq_max + q;
    const unsigned long long t0 = clock64();
    while (ld_acquire_sys(mine) != epoch) {
      if (clock64() - t0 > kSpinTimeoutCycles) {
        *reinterpret_cast<volatile uint32_t*>(status) = 1u + threadIdx.x;  // which source never arrived (host-mapped word)
        __threadfence_system();
        sh_timeout = 1u;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (sh_timeout) {
    // a source never delivered this query: stale bytes must not pass for a result.  The query comes back EMPTY
    // (-1 / -inf) and the status word makes every later call on this exchange fail with AVL_ERR_STATE.
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      out_idx[static_cast<size_t>(q) * k + j] = -1;
      out_val[static_cast<size_t>(q) * k + j] = -INFINITY;
    }
    return;
  }
  // ---- 3. merge world * k entries: (score desc, global id asc), -1 = empty slot
  const int m = v.world * k;
  const uint8_t* base = v.peer[v.rank] + static_cast<uint64_t>(parity) * v.world * v.slot_bytes;
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    const int src = e / k, j = e - src * k;
    const uint8_t* slot = base + static_cast<uint64_t>(src) * v.slot_bytes;
    si[e] = reinterpret_cast<const volatile int64_t*>(slot)[static_cast<size_t>(q) * k + j];
    sv[e] = reinterpret_cast<const volatile float*>(slot + ids_bytes)[static_cast<size_t>(q) * k + j];
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    out_idx[static_cast<size_t>(q) * k + j] = -1;
    out_val[static_cast<size_t>(q) * k + j] = -INFINITY;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    const int64_t i = si[e];
    if (i < 0) continue;
    const float s = sv[e];
    int rank = 0;
    for (int t = 0; t < m; ++t) {
      const int64_t it = si[t];
      const float st = sv[t];
      rank += (it >= 0 && (st > s || (st == s && it < i))) ? 1 : 0;
    }
    if (rank < k) {
      out_idx[static_cast<size_t>(q) * k + rank] = i;
      out_val[static_cast<size_t>(q) * k + rank] = s;
    }
  }
}

}  // namespace
}  // namespace avl

using namespace avl;

struct avl_p2p {
  P2PView view;
  uint8_t* local = nullptr;
  size_t total_bytes = 0;
  uint32_t* status_host = nullptr;  // pinned, device-mapped word: 0 = fine, 1 + source = that source timed out
  uint32_t* status = nullptr;       // its device address (the kernel's atomicExch lands in host memory)
  uint32_t epoch = 0;
  bool connected = false;
  bool opened[kMaxWorld] = {};
};

extern "C" {

int avl_p2p_create(int32_t rank, int32_t world, int32_t nq_max, int32_t k_max, avl_p2p** out) {
  AVL_ARG(out != nullptr, "out is NULL");
  *out = nullptr;
  AVL_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "rank / world out of range (world <= 16)");
  AVL_ARG(nq_max >= 1 && nq_max <= AVL_MAX_QUERIES && k_max >= 1 && k_max <= AVL_MAX_TOPK, "nq_max / k_max out of range");
  AVL_ARG(world * k_max <= 1024, "world * k_max must be <= 1024");
  int dev = 0;
  AVL_CUDA(cudaGetDevice(&dev));
  avl_p2p* p = new avl_p2p();
  memset(&p->view, 0, sizeof(p->view));
  p->view.rank = rank; p->view.world = world; p->view.nq_max = nq_max; p->view.k_max = k_max;
  const uint64_t entries = static_cast<uint64_t>(nq_max) * k_max;
  p->view.slot_bytes = (entries * (sizeof(int64_t) + sizeof(float)) + 255u) & ~uint64_t(255);
  p->view.data_bytes = 2ull * world * p->view.slot_bytes;
  p->total_bytes = p->view.data_bytes + 2ull * world * nq_max * sizeof(uint32_t);
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p->local), p->total_bytes);
  if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->total_bytes);
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&p->status_host), sizeof(uint32_t), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *p->status_host = 0u;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&p->status), p->status_host, 0);
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p->local);
    if (p->status_host) cudaFreeHost(p->status_host);
    delete p;
    return cuda_fail(e, "p2p buffers", __FILE__, __LINE__);
  }
  p->view.peer[rank] = p->local;
  p->connected = (world == 1);
  *out = p;
  return AVL_OK;
}

int avl_p2p_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

int avl_p2p_local_handle(avl_p2p* p, uint8_t* handle) {
  AVL_ARG(p != nullptr && handle != nullptr, "NULL argument");
  cudaIpcMemHandle_t h;
  AVL_CUDA(cudaIpcGetMemHandle(&h, p->local));
  memcpy(handle, &h, sizeof(h));
  return AVL_OK;
}

int avl_p2p_connect(avl_p2p* p, const uint8_t* handles) {
  AVL_ARG(p != nullptr && handles != nullptr, "NULL argument");
  if (p->connected) return AVL_OK;
  for (int r = 0; r < p->view.world; ++r) {
    if (r == p->view.rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + static_cast<size_t>(r) * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    AVL_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->view.peer[r] = static_cast<uint8_t*>(ptr);
    p->opened[r] = true;
  }
  p->connected = true;
  return AVL_OK;
}

int avl_p2p_exchange_merge(avl_p2p* p, const int64_t* idx, const float* val, int32_t nq, int32_t k, int64_t row_offset,
                           const int64_t* global_ids, int64_t* out_idx, float* out_val, int flags, void* stream) {
  AVL_ARG(p != nullptr && idx && val && out_idx && out_val, "NULL argument");
  AVL_ARG(nq >= 1 && nq <= p->view.nq_max && k >= 1 && k <= p->view.k_max, "nq / k exceed what the exchange was created for");
  if (!(flags & AVL_ON_DEVICE)) {
    set_error("avl_p2p_exchange_merge takes device pointers (the slab's top-k as avl_sim_topk left it in HBM)");
    return AVL_ERR_UNSUPPORTED;
  }
  if (!p->connected) {
    set_error("avl_p2p_connect has not been called");
    return AVL_ERR_STATE;
  }
  if (const uint32_t st = *static_cast<volatile uint32_t*>(p->status_host)) {
    // an earlier exchange gave up on a peer (its result was merged from stale data): refuse to go on silently
    char buf[160];
    snprintf(buf, sizeof(buf), "an earlier peer exchange timed out waiting for rank %u; the exchange object is unusable", st - 1u);
    set_error(buf);
    return AVL_ERR_STATE;
  }
  p->epoch += 1;
  p2p_exchange_merge_kernel<<<nq, 128, 0, static_cast<cudaStream_t>(stream)>>>(p->view, idx, val, nq, k, p->epoch, row_offset,
                                                                              global_ids, out_idx, out_val, p->status);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int avl_p2p_status(avl_p2p* p, int32_t* timed_out_source, void* stream) {
  AVL_ARG(p != nullptr && timed_out_source != nullptr, "NULL argument");
  AVL_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  const uint32_t s = *static_cast<volatile uint32_t*>(p->status_host);
  *timed_out_source = s ? static_cast<int32_t>(s) - 1 : -1;
  return AVL_OK;
}

int avl_p2p_destroy(avl_p2p* p) {
  if (!p) return AVL_OK;
  for (int r = 0; r < p->view.world; ++r)
    if (p->opened[r]) cudaIpcCloseMemHandle(p->view.peer[r]);
  cudaFree(p->local);
  if (p->status_host) cudaFreeHost(p->status_host);
  delete p;
  return AVL_OK;
}

}  // extern "C"
