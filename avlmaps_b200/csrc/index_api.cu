// C-ABI of the landmark-index path: host orchestration around the tcgen05 screen and the exact
// kernels.  No torch types, no CPU compute: every result is produced by the kernels in
// sim_screen.cu / sim_exact.cu.
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "avl_internal.h"

namespace avl {

// ------------------------------------------------------------------ errors
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[1024];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e),
           file, line, what);
  g_err = buf;
  return AVL_ERR_CUDA;
}
static bool g_profiling = false;
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("AVL_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// ------------------------------------------------------------------ tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 row-major (rows x dpad) matrix, box = 64 columns (one 128-byte swizzle row) x box_rows.
// The tile-major operand copy of a map is described as (tiles * kblocks * 128) rows of 64 elements: a box is then
// one contiguous run of memory (see map_prepare_kernel).
static int encode_kmajor_map(CUtensorMap* out, const void* base, uint64_t rows, uint64_t dpad,
                             uint32_t box_rows, bool f16 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return AVL_ERR_CUDA;
  }
  cuuint64_t dims[2] = {dpad, rows};
  cuuint64_t strides[1] = {dpad * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", static_cast<int>(r));
    set_error(buf);
    return AVL_ERR_CUDA;
  }
  return AVL_OK;
}

// ------------------------------------------------------------------ workspace
constexpr int kQGlobStride = 4 + AVL_MAX_QUERIES;  // floats of one q_glob half (launch_query_prepare)
struct Workspace {
  float* q = nullptr;          // [256 x d] staged queries (host-pointer calls)
  double* q64 = nullptr;       // [256 x d] queries in fp64 for the exact re-rank
  float* scale = nullptr;      // [256]
  __nv_bfloat16* bq = nullptr; // [256 x dpad]
  float* q_bn = nullptr;       // [256]
  float* q_glob = nullptr;     // [2]
  float* thr_t = nullptr;      // [256]
  uint32_t* flag_count = nullptr;
  uint32_t* flag_rows = nullptr;
  uint32_t* flag_masks = nullptr;
  uint32_t flag_cap = 0;
  uint32_t* cand_cnt = nullptr;  // [256] per-query candidate counters
  uint32_t* cand_row = nullptr;  // [256][cand_cap]
  float* cand_val = nullptr;
  float* cand_val2 = nullptr;    // second per-candidate value (fusion: heat lower bounds)
  uint32_t cand_cap = 0;
  uint32_t* overflow = nullptr;  // [256]
  uint32_t* bucket_cnt = nullptr;  // [num_sms][256] fill counts of the per-CTA candidate buckets (top-k screen)
  void* fb_scratch = nullptr;    // per-block top-k keys of the exact fallback
  size_t fb_scratch_bytes = 0;
  uint32_t* fb_tickets = nullptr;  // [256]
  uint32_t* tile_ctr = nullptr;  // [1] tile counter of the screen kernel's dynamic schedule
  uint32_t tile_base = 0;        // host mirror: the counter's value before the next launch
  // pipelined top-k calls (AVL_PIPELINED): the tail of a call (finalize, fallback, result copy) runs on `tail_stream`
  // next to the following call's screen.  q, scale, q_bn, q_glob, cand_cnt, bucket_cnt, out_idx, out_score and the
  // candidate arrays hold TWO halves, selected by the parity of the call; every other path uses half 0.
  cudaStream_t tail_stream = nullptr;
  cudaEvent_t ev_main[2] = {nullptr, nullptr};   // the screen of the call of this parity is done
  cudaEvent_t ev_tail[2] = {nullptr, nullptr};   // its tail is done (the half's buffers may be reused)
  bool tail_pending[2] = {false, false};
  uint32_t n_pipelined = 0;
  cudaEvent_t tl[2][6] = {};                     // AVL_DEBUG_FLAGS & 128: timeline of a pipelined call (timing events)
  uint32_t* fin_scratch = nullptr;               // [2][256][3 * fin_cap] per-candidate arrays of the finalize beside the screen
  size_t fin_scratch_elems = 0;
  float* sample_t = nullptr;
  size_t sample_elems = 0;
  float* fuse_a = nullptr;       // (pairs, n) dense screen scores of the two modalities (avl_fuse_topk)
  float* fuse_b = nullptr;
  size_t fuse_elems = 0;
  uint32_t* fuse_small = nullptr;  // column statistics, extreme-candidate lists, exact min / max
  int64_t* out_idx = nullptr;  // [256 x AVL_MAX_TOPK]
  float* out_score = nullptr;
  int32_t* argmax = nullptr;   // [n] (host-pointer calls)
  float* column = nullptr;     // [n] scratch column (fallback / fusion)
  void* topk_scratch = nullptr;
  size_t topk_scratch_bytes = 0;
  uint32_t* dbg_host = nullptr;
  uint32_t* dbg_dev = nullptr;
  uint32_t* pin = nullptr;       // pinned host scratch for the small read-backs (pageable copies stage and stall)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ring_ev[2 * 256] = {};  // event pairs around the main screen launch of the last 256 profiled top-k calls
  int ring_n = 0;                      // pairs recorded since the last avl_map_screen_times
};

}  // namespace avl

struct avl_map {
  int device = 0;
  int num_sms = 0;
  int64_t n = 0;
  int32_t d = 0, dpad = 0;
  int op_f16 = 0;         // tensor-core operands are fp16 (AVL_MAP_F16) instead of bf16
  int tiled = 0;          // operand copy is tile-major (map_prepare_kernel)
  float* feat = nullptr;
  __nv_bfloat16* bf = nullptr;
  float* row_norm = nullptr;
  float* row_c = nullptr;
  float* row_an = nullptr;
  CUtensorMap tmap_a;     // box 64 x 128 rows (sim_screen.cu)
  int64_t bytes = 0;
  avl::Workspace ws;
};

namespace avl {

template <typename T>
static int dev_alloc(T** p, size_t count, int64_t* bytes) {
  *p = nullptr;
  if (count == 0) count = 1;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  if (bytes) *bytes += static_cast<int64_t>(count * sizeof(T));
  return AVL_OK;
}

static int ws_init(avl_map* m) {
  Workspace& w = m->ws;
  int rc;
  if ((rc = dev_alloc(&w.q, static_cast<size_t>(2 * AVL_MAX_QUERIES) * m->d, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.q64, static_cast<size_t>(AVL_MAX_QUERIES) * m->d, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.scale, 2 * AVL_MAX_QUERIES, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.bq, static_cast<size_t>(AVL_MAX_QUERIES) * m->dpad, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.q_bn, 2 * AVL_MAX_QUERIES, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.q_glob, 2 * kQGlobStride, &m->bytes))) return rc;  // see launch_query_prepare
  AVL_CUDA(cudaMemset(w.q_glob, 0, 2 * kQGlobStride * sizeof(float)));
  if ((rc = dev_alloc(&w.thr_t, AVL_MAX_QUERIES, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.flag_count, 1, &m->bytes))) return rc;
  // counters and overflow flags are adjacent: one 2 KiB read-back into pinned memory per top-k call
  if ((rc = dev_alloc(&w.cand_cnt, 4 * AVL_MAX_QUERIES, &m->bytes))) return rc;
  w.overflow = w.cand_cnt + AVL_MAX_QUERIES;
  AVL_CUDA(cudaMemset(w.cand_cnt, 0, 4 * AVL_MAX_QUERIES * sizeof(uint32_t)));
  if ((rc = dev_alloc(&w.bucket_cnt, static_cast<size_t>(2 * 256) * AVL_MAX_QUERIES, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.fb_tickets, AVL_MAX_QUERIES, &m->bytes))) return rc;
  AVL_CUDA(cudaMemset(w.fb_tickets, 0, AVL_MAX_QUERIES * sizeof(uint32_t)));
  if ((rc = dev_alloc(&w.tile_ctr, 1, &m->bytes))) return rc;
  AVL_CUDA(cudaMemset(w.tile_ctr, 0, sizeof(uint32_t)));
  AVL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&w.pin), 2 * AVL_MAX_QUERIES * sizeof(uint32_t), cudaHostAllocDefault));
  if ((rc = dev_alloc(&w.out_idx, static_cast<size_t>(2 * AVL_MAX_QUERIES) * AVL_MAX_TOPK, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.out_score, static_cast<size_t>(2 * AVL_MAX_QUERIES) * AVL_MAX_TOPK, &m->bytes))) return rc;
  {
    // lowest priority: the tail fills what the next call's head and screen leave free, it must never delay them
    int lo = 0, hi = 0;
    AVL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    AVL_CUDA(cudaStreamCreateWithPriority(&w.tail_stream, cudaStreamNonBlocking, lo));
  }
  for (int i = 0; i < 2; ++i) {
    AVL_CUDA(cudaEventCreateWithFlags(&w.ev_main[i], cudaEventDisableTiming));
    AVL_CUDA(cudaEventCreateWithFlags(&w.ev_tail[i], cudaEventDisableTiming));
  }
  AVL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&w.dbg_host), 4096, cudaHostAllocMapped));
  memset(w.dbg_host, 0, 4096);
  AVL_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void**>(&w.dbg_dev), w.dbg_host, 0));
  for (int i = 0; i < 4; ++i) AVL_CUDA(cudaEventCreate(&w.ev[i]));
  for (int i = 0; i < 2 * 256; ++i) AVL_CUDA(cudaEventCreate(&w.ring_ev[i]));
  return AVL_OK;
}

static void ws_free(Workspace& w) {
  cudaFree(w.q); cudaFree(w.q64); cudaFree(w.scale); cudaFree(w.bq); cudaFree(w.q_bn); cudaFree(w.q_glob); cudaFree(w.thr_t);
  cudaFree(w.flag_count); cudaFree(w.flag_rows); cudaFree(w.flag_masks); cudaFree(w.cand_cnt);
  cudaFree(w.cand_row); cudaFree(w.cand_val); cudaFree(w.sample_t); cudaFree(w.out_idx);
  cudaFree(w.out_score); cudaFree(w.argmax); cudaFree(w.column); cudaFree(w.topk_scratch);
  cudaFree(w.fuse_a); cudaFree(w.fuse_b); cudaFree(w.fuse_small); cudaFree(w.cand_val2);
  cudaFree(w.bucket_cnt); cudaFree(w.fb_scratch); cudaFree(w.fb_tickets); cudaFree(w.tile_ctr); cudaFree(w.fin_scratch);
  for (int i = 0; i < 2; ++i) {
    if (w.ev_main[i]) cudaEventDestroy(w.ev_main[i]);
    if (w.ev_tail[i]) cudaEventDestroy(w.ev_tail[i]);
  }
  if (w.tail_stream) cudaStreamDestroy(w.tail_stream);
  if (w.dbg_host) cudaFreeHost(w.dbg_host);
  if (w.pin) cudaFreeHost(w.pin);
  for (int i = 0; i < 4; ++i)
    if (w.ev[i]) cudaEventDestroy(w.ev[i]);
  for (int i = 0; i < 2 * 256; ++i)
    if (w.ring_ev[i]) cudaEventDestroy(w.ring_ev[i]);
}

static int ensure_flags(avl_map* m) {
  Workspace& w = m->ws;
  const uint32_t need = static_cast<uint32_t>(std::max<int64_t>(m->n, 1));
  if (w.flag_cap >= need) return AVL_OK;
  cudaFree(w.flag_rows); cudaFree(w.flag_masks);
  int rc;
  if ((rc = dev_alloc(&w.flag_rows, need, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.flag_masks, static_cast<size_t>(need) * kFlagWords, &m->bytes))) return rc;
  w.flag_cap = need;
  return AVL_OK;
}
static int ensure_cands(avl_map* m, uint32_t cap) {
  Workspace& w = m->ws;
  if (w.cand_cap >= cap) return AVL_OK;
  cudaFree(w.cand_row); cudaFree(w.cand_val); cudaFree(w.cand_val2);
  w.cand_val2 = nullptr;
  int rc;
  if ((rc = dev_alloc(&w.cand_row, static_cast<size_t>(AVL_MAX_QUERIES) * cap, &m->bytes))) return rc;
  if ((rc = dev_alloc(&w.cand_val, static_cast<size_t>(AVL_MAX_QUERIES) * cap, &m->bytes))) return rc;
  w.cand_cap = cap;
  return AVL_OK;
}
static int ensure_fallback(avl_map* m, int32_t k) {
  Workspace& w = m->ws;
  const size_t need = static_cast<size_t>(AVL_MAX_QUERIES) * (2 * m->num_sms) * k * sizeof(unsigned long long);
  if (w.fb_scratch_bytes >= need) return AVL_OK;
  cudaFree(w.fb_scratch);
  w.fb_scratch = nullptr;
  w.fb_scratch_bytes = 0;
  AVL_CUDA(cudaMalloc(&w.fb_scratch, need));
  w.fb_scratch_bytes = need;
  m->bytes += static_cast<int64_t>(need);
  return AVL_OK;
}
static int ensure_sample(avl_map* m, size_t elems) {
  Workspace& w = m->ws;
  if (w.sample_elems >= elems) return AVL_OK;
  cudaFree(w.sample_t);
  int rc;
  if ((rc = dev_alloc(&w.sample_t, elems, &m->bytes))) return rc;
  w.sample_elems = elems;
  return AVL_OK;
}
static int ensure_column(avl_map* m) {
  Workspace& w = m->ws;
  int rc;
  if (!w.column && (rc = dev_alloc(&w.column, static_cast<size_t>(std::max<int64_t>(m->n, 1)), &m->bytes)))
    return rc;
  const size_t need = topk_vector_scratch_bytes(m->n);
  if (w.topk_scratch_bytes < need) {
    cudaFree(w.topk_scratch);
    AVL_CUDA(cudaMalloc(&w.topk_scratch, need));
    w.topk_scratch_bytes = need;
    m->bytes += static_cast<int64_t>(need);
  }
  return AVL_OK;
}

static int check_watchdog(avl_map* m, int rc) {
  if (rc != AVL_OK && m->ws.dbg_host && (m->ws.dbg_host[0] >> 16) == 0xDEADu) {
    char buf[256];
    snprintf(buf, sizeof(buf), " | screen watchdog: tag=0x%x block=%u thread=%u parity=%u",
             m->ws.dbg_host[0] & 0xFFFFu, m->ws.dbg_host[1], m->ws.dbg_host[2], m->ws.dbg_host[3]);
    g_err += buf;
  }
  return rc;
}

// which tcgen05 variant: 2 when B only fits the shared memory of an SM pair (or forced by env)
static int pick_cta_group(int npad, int kblocks, int forced, int mode) {
  if (forced == 1 || forced == 2) return screen_pick_stages(forced, npad, kblocks, mode) >= 2 ? forced : 0;
  const char* env = getenv("AVL_CTA_GROUP");
  if (env && (env[0] == '1' || env[0] == '2')) {
    const int f = env[0] - '0';
    if (screen_pick_stages(f, npad, kblocks, mode) >= 2) return f;
  }
  if (screen_pick_stages(1, npad, kblocks, mode) >= 4) return 1;
  if (screen_pick_stages(2, npad, kblocks, mode) >= 2) return 2;
  if (screen_pick_stages(1, npad, kblocks, mode) >= 2) return 1;
  return 0;
}

struct QuerySetup {
  const float* q_dev = nullptr;      // fp32 queries on device
  const float* scale_dev = nullptr;  // or null
  int npad = 0;
  int cg = 0;
  float* q_bn = nullptr;             // per-query ||b|| and the two maxima: the workspace half of this call's parity
  float* q_glob = nullptr;
  int unit_rows = 0;                 // voxel rows per tile unit of the chosen kernel
  int stages = 0;
  size_t smem = 0;
  CUtensorMap tmap_b;
};

// stage queries, build bf16 B + norms, pick the kernel variant
static int setup_queries(avl_map* m, const float* queries, int32_t nq, const float* scale, int flags,
                         int forced_cg, int screen_mode /* ScreenMode of the pass that follows, -1: no screen */,
                         cudaStream_t s, QuerySetup* qs, bool fold_scale = false, int parity = 0) {
  Workspace& w = m->ws;
  AVL_ARG(queries != nullptr, "queries is NULL");
  AVL_ARG(nq >= 1 && nq <= AVL_MAX_QUERIES, "nq must be in [1, AVL_MAX_QUERIES]");
  qs->q_bn = w.q_bn + parity * AVL_MAX_QUERIES;
  qs->q_glob = w.q_glob + parity * kQGlobStride;
  if (flags & AVL_ON_DEVICE) {
    qs->q_dev = queries;
    qs->scale_dev = scale;
  } else {
    float* qd = w.q + static_cast<size_t>(parity) * AVL_MAX_QUERIES * m->d;
    AVL_CUDA(cudaMemcpyAsync(qd, queries, static_cast<size_t>(nq) * m->d * sizeof(float), cudaMemcpyHostToDevice, s));
    qs->q_dev = qd;
    if (scale) {
      float* sd = w.scale + parity * AVL_MAX_QUERIES;
      AVL_CUDA(cudaMemcpyAsync(sd, scale, static_cast<size_t>(nq) * sizeof(float), cudaMemcpyHostToDevice, s));
      qs->scale_dev = sd;
    }
  }
  if (screen_mode < 0) return AVL_OK;
  qs->npad = (nq + 15) & ~15;
  const int kblocks = m->dpad / kBlockK;
  qs->cg = pick_cta_group(qs->npad, kblocks, forced_cg, screen_mode);
  if (qs->cg == 0) {
    set_error("query batch does not fit the shared memory of an SM pair (nq * dim too large); split the batch");
    return AVL_ERR_UNSUPPORTED;
  }
  qs->unit_rows = kTileRows * qs->cg;
  qs->stages = screen_pick_stages(qs->cg, qs->npad, kblocks, screen_mode);
  qs->smem = screen_smem_bytes(qs->cg, qs->npad, kblocks, qs->stages, screen_mode);
  int rc = launch_query_prepare(qs->q_dev, fold_scale ? qs->scale_dev : nullptr, nq, m->d, m->dpad, qs->npad, w.bq,
                                qs->q_bn, qs->q_glob, m->op_f16, s);
  if (rc) return rc;
  return encode_kmajor_map(&qs->tmap_b, w.bq, static_cast<uint64_t>(qs->npad), static_cast<uint64_t>(m->dpad),
                           static_cast<uint32_t>(qs->npad / qs->cg), m->op_f16 != 0);
}

static void base_params(const avl_map* m, const QuerySetup& qs, int32_t nq, int normalize, ScreenParams* p) {
  memset(p, 0, sizeof(*p));
  p->n_rows = m->n;
  p->kblocks = m->dpad / kBlockK;
  p->nq = nq;
  p->npad = qs.npad;
  p->stages = qs.stages;
  p->normalize = normalize;
  p->row_norm = m->row_norm;
  p->row_c = m->row_c;
  p->row_an = m->row_an;
  p->q_bn = qs.q_bn;
  p->bq = m->ws.bq;
  p->q_glob = qs.q_glob;
  p->dbg = m->ws.dbg_dev;
  p->op_f16 = m->op_f16;
  p->a_tiled = m->tiled;
  p->a_base = m->bf;
  p->tile_stride = 1;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e2 = getenv("AVL_DEBUG_FLAGS");
      dbg = e2 ? atoi(e2) : 0;
    }
    p->debug_flags = dbg;
  }
  const int64_t unit = qs.unit_rows;
  p->num_tiles = static_cast<int32_t>((m->n + unit - 1) / unit);
}

static int run_screen(avl_map* m, const QuerySetup& qs, ScreenParams& p, cudaStream_t s) {
  // dynamic tile schedule: one global counter per map, never reset; a launch makes exactly num_tiles fetches
  p.tile_ctr = m->ws.tile_ctr;
  p.tile_base = m->ws.tile_base;
  m->ws.tile_base += static_cast<uint32_t>(p.num_tiles);
  return launch_screen(qs.cg, &m->tmap_a, &qs.tmap_b, p, m->num_sms, qs.smem, s);
}

// Order `s` after the tails of earlier pipelined top-k calls (they run on the map's tail stream and use its workspace).
static int drain_tails(avl_map* m, cudaStream_t s) {
  Workspace& w = m->ws;
  for (int i = 0; i < 2; ++i) {
    if (w.tail_pending[i]) {
      AVL_CUDA(cudaStreamWaitEvent(s, w.ev_tail[i], 0));
      w.tail_pending[i] = false;
    }
  }
  return AVL_OK;
}

static int scale_positive(const float* scale, int32_t nq, int flags) {
  // the threshold screen divides by scale; the host-pointer path can check it, the device path trusts it
  if (scale && !(flags & AVL_ON_DEVICE))
    for (int i = 0; i < nq; ++i)
      if (!(scale[i] > 0.f) || !std::isfinite(scale[i])) {
        set_error("scale entries must be finite and > 0");
        return AVL_ERR_ARG;
      }
  return AVL_OK;
}

}  // namespace avl

using namespace avl;

extern "C" {

int avl_version(void) { return 100; }
const char* avl_last_error(void) { return g_err.c_str(); }

int avl_device_count(int* count) {
  AVL_ARG(count != nullptr, "count is NULL");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    cudaGetLastError();
    c = 0;
  }
  *count = c;
  return AVL_OK;
}
int avl_set_device(int device) {
  AVL_CUDA(cudaSetDevice(device));
  return AVL_OK;
}
int avl_set_profiling(int enabled) {
  g_profiling = enabled != 0;
  return AVL_OK;
}

int avl_map_create(const float* grid_feat, int64_t n, int32_t dim, int flags, void* stream, avl_map** out) {
  AVL_ARG(out != nullptr, "out is NULL");
  *out = nullptr;
  AVL_ARG(n >= 0 && n < (int64_t(1) << 31), "n must be in [0, 2^31)");
  AVL_ARG(dim >= 1 && dim <= 4096, "dim must be in [1, 4096]");
  AVL_ARG(n == 0 || grid_feat != nullptr, "grid_feat is NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int device = 0, num_sms = 0, major = 0;  // queried before the handle exists: a box without a GPU fails here
  AVL_CUDA(cudaGetDevice(&device));
  AVL_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
  AVL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10) {
    set_error("avlmaps_b200 needs an sm_100a (B200) device");
    return AVL_ERR_UNSUPPORTED;
  }
  avl_map* m = new avl_map();
  m->device = device;
  m->num_sms = num_sms;
  m->n = n;
  m->d = dim;
  m->dpad = (dim + kBlockK - 1) / kBlockK * kBlockK;
  int rc = AVL_OK;
  const size_t rows = static_cast<size_t>(std::max<int64_t>(n, 1));
  do {
    if ((rc = dev_alloc(&m->feat, rows * dim, &m->bytes))) break;
    // operand copy: whole 128-row tiles (the tile-major layout addresses by tile; the tail rows are zero)
    const size_t tile_rows_total = (rows + kTileRows - 1) / kTileRows * kTileRows;
    {
      const char* e = getenv("AVL_TILED");
      m->tiled = !(e && e[0] == '0');
    }
    if ((rc = dev_alloc(&m->bf, tile_rows_total * m->dpad, &m->bytes))) break;
    {
      const size_t tail = static_cast<size_t>(kTileRows) * m->dpad;  // last tile: rows past n stay zero
      cudaError_t e = cudaMemsetAsync(m->bf + (tile_rows_total * m->dpad - tail), 0, tail * sizeof(__nv_bfloat16), s);
      if (e != cudaSuccess) { rc = cuda_fail(e, "clear operand tail", __FILE__, __LINE__); break; }
    }
    if ((rc = dev_alloc(&m->row_norm, rows, &m->bytes))) break;
    if ((rc = dev_alloc(&m->row_c, rows, &m->bytes))) break;
    if ((rc = dev_alloc(&m->row_an, rows, &m->bytes))) break;
    if ((rc = ws_init(m))) break;
    if (n > 0) {
      cudaError_t e = cudaMemcpyAsync(m->feat, grid_feat, static_cast<size_t>(n) * dim * sizeof(float),
                                      (flags & AVL_ON_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) { rc = cuda_fail(e, "upload grid_feat", __FILE__, __LINE__); break; }
    }
    // kappa: fp32 accumulation slack of the tensor core over dpad products
    const float kappa = static_cast<float>(m->dpad) * 2.4e-7f;
    m->op_f16 = (flags & AVL_MAP_F16) ? 1 : 0;
    uint32_t* nonfinite = m->ws.flag_count;  // scratch word of the workspace
    for (int attempt = 0; attempt < 2; ++attempt) {
      uint32_t bad = 0;
      cudaError_t e = cudaMemsetAsync(nonfinite, 0, sizeof(uint32_t), s);
      if (e != cudaSuccess) { rc = cuda_fail(e, "map_prepare", __FILE__, __LINE__); break; }
      if ((rc = launch_map_prepare(m->feat, n, dim, m->dpad, m->bf, m->row_norm, m->row_c, m->row_an, kappa, m->op_f16,
                                   nonfinite, m->tiled, s))) break;
      e = cudaMemcpyAsync(&bad, nonfinite, sizeof(bad), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { rc = cuda_fail(e, "map_prepare", __FILE__, __LINE__); break; }
      if (!(bad && m->op_f16)) break;
      m->op_f16 = 0;  // a value beyond the fp16 range: bf16 operands keep the fp32 range
    }
    if (rc) break;
    const uint64_t map_rows = m->tiled ? static_cast<uint64_t>(tile_rows_total) * (m->dpad / kBlockK) : rows;
    const uint64_t map_cols = m->tiled ? static_cast<uint64_t>(kBlockK) : static_cast<uint64_t>(m->dpad);
    if ((rc = encode_kmajor_map(&m->tmap_a, m->bf, map_rows, map_cols, kTileRows, m->op_f16 != 0))) break;
  } while (0);
  if (rc) {
    avl_map_destroy(m);
    return rc;
  }
  *out = m;
  return AVL_OK;
}

int avl_map_destroy(avl_map* m) {
  if (!m) return AVL_OK;
  if (m->ws.tail_stream) cudaStreamSynchronize(m->ws.tail_stream);
  ws_free(m->ws);
  cudaFree(m->feat); cudaFree(m->bf); cudaFree(m->row_norm); cudaFree(m->row_c); cudaFree(m->row_an);
  delete m;
  return AVL_OK;
}

int avl_map_operand_f16(const avl_map* m) { return m ? m->op_f16 : 0; }

int avl_map_shape(const avl_map* m, int64_t* n, int32_t* dim) {
  AVL_ARG(m != nullptr, "map is NULL");
  if (n) *n = m->n;
  if (dim) *dim = m->d;
  return AVL_OK;
}
int64_t avl_map_device_bytes(const avl_map* m) { return m ? m->bytes : 0; }

int avl_sim_dense(avl_map* m, const float* queries, int32_t nq, const float* scale, int normalize_map,
                  float* out_scores, int flags, void* stream) {
  AVL_ARG(m != nullptr && out_scores != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  QuerySetup qs;
  int rc = drain_tails(m, s);
  if (rc) return rc;
  if ((rc = setup_queries(m, queries, nq, scale, flags, 0, -1, s, &qs))) return rc;
  if (m->n == 0) return AVL_OK;
  if (flags & AVL_ON_DEVICE)
    return launch_dense_exact(m->feat, m->n, m->d, qs.q_dev, nq, qs.scale_dev, m->row_norm, normalize_map,
                              out_scores, nq, 1, s);
  // host output: row chunks through a bounded device buffer
  const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(m->n, (int64_t(256) << 20) / (4 * nq)));
  float* buf = nullptr;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&buf), static_cast<size_t>(chunk) * nq * sizeof(float)));
  for (int64_t r0 = 0; r0 < m->n && rc == AVL_OK; r0 += chunk) {
    const int64_t rows = std::min(chunk, m->n - r0);
    rc = launch_dense_exact(m->feat + r0 * m->d, rows, m->d, qs.q_dev, nq, qs.scale_dev, m->row_norm + r0,
                            normalize_map, buf, nq, 1, s);
    if (rc) break;
    cudaError_t e = cudaMemcpyAsync(out_scores + r0 * nq, buf, static_cast<size_t>(rows) * nq * sizeof(float),
                                    cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = cuda_fail(e, "dense D2H", __FILE__, __LINE__);
  }
  cudaFree(buf);
  return rc;
}

int avl_sim_screen_dense(avl_map* m, const float* queries, int32_t nq, int32_t cta_group, float* out_scores,
                         int flags, void* stream) {
  AVL_ARG(m != nullptr && out_scores != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  QuerySetup qs;
  int rc = drain_tails(m, s);
  if (rc) return rc;
  if ((rc = setup_queries(m, queries, nq, nullptr, flags, cta_group, kModeDense, s, &qs))) return rc;
  if (m->n == 0) return AVL_OK;
  float* dst = out_scores;
  if (!(flags & AVL_ON_DEVICE))
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dst), static_cast<size_t>(m->n) * nq * sizeof(float)));
  ScreenParams p;
  base_params(m, qs, nq, 0, &p);
  p.mode = kModeDense;
  p.dense_out = dst;
  p.dense_rs = nq;
  p.dense_cs = 1;
  p.dense_cols = nq;
  rc = run_screen(m, qs, p, s);
  if (rc == AVL_OK && !(flags & AVL_ON_DEVICE)) {
    cudaError_t e = cudaMemcpyAsync(out_scores, dst, static_cast<size_t>(m->n) * nq * sizeof(float),
                                    cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) rc = cuda_fail(e, "screen_dense", __FILE__, __LINE__);
  }
  if (!(flags & AVL_ON_DEVICE)) cudaFree(dst);
  return check_watchdog(m, rc);
}

int avl_sim_argmax(avl_map* m, const float* queries, int32_t nq, const float* scale, int normalize_map,
                   int32_t* out_argmax, int flags, void* stream, avl_index_stats* stats) {
  AVL_ARG(m != nullptr && out_argmax != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace& w = m->ws;
  int rc = scale_positive(scale, nq, flags);
  if (rc) return rc;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->n_rows = m->n;
    stats->dim = m->d;
    stats->n_queries = nq;
  }
  if (m->n == 0) return AVL_OK;
  if ((rc = drain_tails(m, s))) return rc;
  if (g_profiling) AVL_CUDA(cudaEventRecord(w.ev[0], s));
  QuerySetup qs;
  if ((rc = setup_queries(m, queries, nq, scale, flags, 0, kModeArgmax, s, &qs, /*fold_scale=*/true))) return rc;
  if ((rc = ensure_flags(m))) return rc;
  int32_t* dst = out_argmax;
  if (!(flags & AVL_ON_DEVICE)) {
    if (!w.argmax && (rc = dev_alloc(&w.argmax, static_cast<size_t>(m->n), &m->bytes))) return rc;
    dst = w.argmax;
  }
  AVL_CUDA(cudaMemsetAsync(w.flag_count, 0, sizeof(uint32_t), s));
  ScreenParams p;
  base_params(m, qs, nq, normalize_map, &p);
  p.mode = kModeArgmax;
  p.argmax_out = dst;
  p.flag_count = w.flag_count;
  p.flag_rows = w.flag_rows;
  p.flag_masks = w.flag_masks;
  p.flag_cap = w.flag_cap;
  if (g_profiling) AVL_CUDA(cudaEventRecord(w.ev[1], s));
  if ((rc = run_screen(m, qs, p, s))) return rc;
  if (g_profiling) AVL_CUDA(cudaEventRecord(w.ev[2], s));
  if ((rc = launch_argmax_rerank(m->feat, m->d, qs.q_dev, w.q64, w.q_bn, nq, qs.scale_dev, m->row_norm, normalize_map,
                                 w.flag_count, w.flag_rows, w.flag_masks, w.flag_cap, dst, m->num_sms, s)))
    return rc;
  if (g_profiling) AVL_CUDA(cudaEventRecord(w.ev[3], s));
  if (!(flags & AVL_ON_DEVICE))
    AVL_CUDA(cudaMemcpyAsync(out_argmax, dst, static_cast<size_t>(m->n) * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (stats || !(flags & AVL_ON_DEVICE) || g_profiling) {
    uint32_t nflag = 0;
    cudaError_t e = cudaMemcpyAsync(&nflag, w.flag_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return check_watchdog(m, cuda_fail(e, "argmax", __FILE__, __LINE__));
    if (stats) {
      stats->cta_group = qs.cg;
      stats->n_launches = 4;  // query_prepare, screen, f32->f64, rerank
      stats->n_flagged = nflag;
      if (g_profiling) {
        cudaEventElapsedTime(&stats->ms_screen, w.ev[1], w.ev[2]);
        cudaEventElapsedTime(&stats->ms_total, w.ev[0], w.ev[3]);
      }
    }
  }
  return AVL_OK;
}

int avl_topk_f32(const float* values, int64_t n, int32_t k, int64_t* out_idx, float* out_val, int flags,
                 void* stream) {
  AVL_ARG(out_idx != nullptr && out_val != nullptr, "NULL output");
  AVL_ARG(k >= 1 && k <= AVL_MAX_TOPK, "k must be in [1, AVL_MAX_TOPK]");
  AVL_ARG(n >= 0 && n < (int64_t(1) << 32), "n out of range");
  AVL_ARG(n == 0 || values != nullptr, "values is NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t sb = topk_vector_scratch_bytes(n);
  void* scratch = nullptr;
  float* dv = nullptr;
  int64_t* di = nullptr;
  float* dval = nullptr;
  int rc = AVL_OK;
  cudaError_t e = cudaMalloc(&scratch, sb);
  if (e != cudaSuccess) return cuda_fail(e, "topk scratch", __FILE__, __LINE__);
  const float* src = values;
  if (!(flags & AVL_ON_DEVICE)) {
    if (cudaMalloc(reinterpret_cast<void**>(&dv), std::max<size_t>(1, n) * sizeof(float)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&di), k * sizeof(int64_t)) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&dval), k * sizeof(float)) != cudaSuccess) {
      rc = cuda_fail(cudaGetLastError(), "topk staging", __FILE__, __LINE__);
    } else {
      e = cudaMemcpyAsync(dv, values, static_cast<size_t>(n) * sizeof(float), cudaMemcpyHostToDevice, s);
      if (e != cudaSuccess) rc = cuda_fail(e, "topk H2D", __FILE__, __LINE__);
      src = dv;
    }
  }
  if (rc == AVL_OK)
    rc = launch_topk_vector(src, n, k, (flags & AVL_ON_DEVICE) ? out_idx : di,
                            (flags & AVL_ON_DEVICE) ? out_val : dval, scratch, sb, s);
  if (rc == AVL_OK && !(flags & AVL_ON_DEVICE)) {
    e = cudaMemcpyAsync(out_idx, di, k * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_val, dval, k * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) rc = cuda_fail(e, "topk D2H", __FILE__, __LINE__);
  }
  e = cudaStreamSynchronize(s);  // scratch is freed below
  if (e != cudaSuccess && rc == AVL_OK) rc = cuda_fail(e, "topk", __FILE__, __LINE__);
  cudaFree(scratch); cudaFree(dv); cudaFree(di); cudaFree(dval);
  return rc;
}

int avl_sim_topk(avl_map* m, const float* queries, int32_t nq, const float* scale, int normalize_map,
                 int32_t k, int64_t* out_idx, float* out_score, int flags, void* stream,
                 avl_index_stats* stats) {
  AVL_ARG(m != nullptr && out_idx != nullptr && out_score != nullptr, "NULL argument");
  AVL_ARG(k >= 1 && k <= AVL_MAX_TOPK, "k must be in [1, AVL_MAX_TOPK]");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  Workspace& w = m->ws;
  int rc = scale_positive(scale, nq, flags);
  if (rc) return rc;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->n_rows = m->n;
    stats->dim = m->d;
    stats->n_queries = nq;
  }
  // Pipelined form (AVL_PIPELINED, no stats): the tail of this call -- finalize, fallback, result copy -- goes to the
  // map's tail stream and runs NEXT TO the following call's screen (the finalize keeps its arrays in global memory and
  // fits beside a screen CTA).  Buffers alternate by the parity of the call; a half is reused only after its tail.
  const bool pipelined = (flags & AVL_PIPELINED) && !stats && m->n > 0;
  const int par = pipelined ? static_cast<int>(w.n_pipelined & 1u) : 0;
  if (pipelined) {
    if (w.tail_pending[par]) {   // the call two back used this half
      AVL_CUDA(cudaStreamWaitEvent(s, w.ev_tail[par], 0));
      w.tail_pending[par] = false;
    }
  } else if ((rc = drain_tails(m, s))) {
    return rc;
  }
  if (g_profiling) AVL_CUDA(cudaEventRecord(w.ev[0], s));
  static const bool timeline = [] { const char* e = getenv("AVL_DEBUG_FLAGS"); return e && (atoi(e) & 128); }();
  if (timeline && pipelined) {
    for (int i = 0; i < 6; ++i)
      if (!w.tl[par][i]) AVL_CUDA(cudaEventCreate(&w.tl[par][i]));
    if (w.n_pipelined >= 2 && (w.n_pipelined % 7) == 2) {   // print the call two back (same parity), now complete
      cudaEventSynchronize(w.tl[par][5]);
      float t[6] = {0};
      for (int i = 1; i < 6; ++i) cudaEventElapsedTime(&t[i], w.tl[par][0], w.tl[par][i]);
      float gap = 0.f;
      if (w.tl[par ^ 1][0]) cudaEventElapsedTime(&gap, w.tl[par][0], w.tl[par ^ 1][0]);
      fprintf(stderr, "[avl timeline] head 0 | main start %.3f | main end %.3f | tail start %.3f | finalize end %.3f | tail end %.3f | next call's head starts %.3f ms\n",
              t[1], t[2], t[3], t[4], t[5], gap);
    }
    AVL_CUDA(cudaEventRecord(w.tl[par][0], s));
  }
  QuerySetup qs;
  if ((rc = setup_queries(m, queries, nq, scale, flags, 0, kModeThresh, s, &qs, false, par))) return rc;
  const size_t out_half = static_cast<size_t>(par) * AVL_MAX_QUERIES * AVL_MAX_TOPK;
  int64_t* d_idx = (flags & AVL_ON_DEVICE) ? out_idx : w.out_idx + out_half;
  float* d_score = (flags & AVL_ON_DEVICE) ? out_score : w.out_score + out_half;

  const int64_t unit = qs.unit_rows;
  const int64_t total_units = (m->n + unit - 1) / unit;
  // rows sampled for the threshold: ~n/64, enough that ~k*n/n0 candidates per query stay << the finalize capacity
  int64_t n0 = std::max<int64_t>(m->n / 64, static_cast<int64_t>(k) * m->n / 1024);
  n0 = std::min<int64_t>(std::max<int64_t>(n0, 8192), std::max<int64_t>(m->n, 1));
  int64_t sample_units = std::min<int64_t>(total_units, (n0 + unit - 1) / unit);
  int64_t tile_stride = sample_units > 0 ? std::max<int64_t>(1, total_units / sample_units) : 1;
  if (sample_units > 0) sample_units = std::min<int64_t>(sample_units, (total_units + tile_stride - 1) / tile_stride);
  const int64_t n_sample = sample_units * unit;
  // candidates: one bucket per (query, CTA of the screen launch); the finalize block of a query holds fin_cap of them
  const uint32_t fin_cap = 8192;
  const int grid = screen_grid(qs.cg, m->num_sms, static_cast<int>(std::min<int64_t>(total_units, 1 << 30)));
  const uint32_t bucket = std::min<uint32_t>(fin_cap, std::max<uint32_t>(128u, 4u * fin_cap / static_cast<uint32_t>(std::max(grid, 1))));
  const uint32_t cand_stride = static_cast<uint32_t>(std::max(grid, 1)) * bucket;   // entries per query
  if ((rc = ensure_cands(m, pipelined ? 2 * cand_stride : cand_stride))) return rc;
  if ((rc = ensure_sample(m, static_cast<size_t>(std::max<int64_t>(n_sample, 1)) * nq))) return rc;
  if ((rc = ensure_fallback(m, k))) return rc;
  if (pipelined && w.fin_scratch_elems < static_cast<size_t>(2) * AVL_MAX_QUERIES * 3 * fin_cap) {
    if ((rc = drain_tails(m, s))) return rc;
    AVL_CUDA(cudaStreamSynchronize(s));
    cudaFree(w.fin_scratch);
    w.fin_scratch = nullptr;
    w.fin_scratch_elems = 0;
    if ((rc = dev_alloc(&w.fin_scratch, static_cast<size_t>(2) * AVL_MAX_QUERIES * 3 * fin_cap, &m->bytes))) return rc;
    w.fin_scratch_elems = static_cast<size_t>(2) * AVL_MAX_QUERIES * 3 * fin_cap;
  }
  uint32_t* cand_row = w.cand_row + static_cast<size_t>(par) * AVL_MAX_QUERIES * cand_stride;
  float* cand_val = w.cand_val + static_cast<size_t>(par) * AVL_MAX_QUERIES * cand_stride;
  uint32_t* bucket_cnt = w.bucket_cnt + static_cast<size_t>(par) * 256 * AVL_MAX_QUERIES;
  uint32_t* cand_cnt = w.cand_cnt + par * 2 * AVL_MAX_QUERIES;
  uint32_t* overflow = cand_cnt + AVL_MAX_QUERIES;

  ScreenParams p;
  if (m->n > 0) {
    // phase A: lower bounds of the sampled tiles, reduced in the screen's epilogue to one maximum per 32-row group
    // and query (dense_lb = 2)
    const int64_t n_cols = n_sample / 32;
    base_params(m, qs, nq, normalize_map, &p);
    p.mode = kModeDense;
    p.num_tiles = static_cast<int32_t>(sample_units);
    p.tile_stride = static_cast<int32_t>(tile_stride);
    p.dense_out = w.sample_t;
    p.dense_rs = 1;
    p.dense_cs = n_cols;
    p.dense_cols = nq;
    p.dense_lb = 2;
    if ((rc = run_screen(m, qs, p, s))) return rc;
    if ((rc = launch_select_threshold(w.sample_t, static_cast<int32_t>(n_cols), n_cols, nq, k, w.thr_t, s)))
      return rc;
    // phase B: full pass, candidates = rows whose upper bound reaches the threshold
    base_params(m, qs, nq, normalize_map, &p);
    p.mode = kModeThresh;
    p.thr_t = w.thr_t;
    p.cand_cnt = bucket_cnt;
    p.cand_row = cand_row;
    p.cand_val = cand_val;
    p.cand_bucket = bucket;
    const int slot = w.ring_n % 256;
    if (timeline && pipelined) AVL_CUDA(cudaEventRecord(w.tl[par][1], s));
    if (g_profiling) {
      AVL_CUDA(cudaEventRecord(w.ev[1], s));
      AVL_CUDA(cudaEventRecord(w.ring_ev[2 * slot], s));
    }
    if ((rc = run_screen(m, qs, p, s))) return rc;
    if (g_profiling) {
      AVL_CUDA(cudaEventRecord(w.ev[2], s));
      AVL_CUDA(cudaEventRecord(w.ring_ev[2 * slot + 1], s));
      ++w.ring_n;
    }
  }
  // phase C: exact re-score of the survivors, final order; totals and overflow flags per query
  // phase D: queries whose buckets overflowed (adversarial data, massive ties) are re-scored exactly -- decided on
  // the device: the kernel leaves at once when no flag is set, so the call needs no host round trip
  cudaStream_t ts = s;   // stream of the tail
  if (pipelined) {
    ts = w.tail_stream;
    if (timeline) AVL_CUDA(cudaEventRecord(w.tl[par][2], s));
    AVL_CUDA(cudaEventRecord(w.ev_main[par], s));
    AVL_CUDA(cudaStreamWaitEvent(ts, w.ev_main[par], 0));
    if (timeline) AVL_CUDA(cudaEventRecord(w.tl[par][3], ts));
  }
  uint32_t* gscratch = pipelined ? w.fin_scratch + static_cast<size_t>(par) * AVL_MAX_QUERIES * 3 * fin_cap : nullptr;
  if ((rc = launch_topk_finalize(m->feat, m->n, m->d, qs.q_dev, nq, qs.scale_dev, m->row_norm, m->row_c,
                                 m->row_an, qs.q_bn, qs.q_glob, normalize_map, k, bucket_cnt, m->n > 0 ? grid : 0,
                                 bucket, cand_row, cand_val, fin_cap, d_idx, d_score, cand_cnt, overflow, gscratch, ts)))
    return rc;
  if (timeline && pipelined) AVL_CUDA(cudaEventRecord(w.tl[par][4], ts));
  if ((rc = launch_topk_fallback(m->feat, m->n, m->d, qs.q_dev, nq, qs.scale_dev, m->row_norm, normalize_map, k,
                                 overflow, w.fb_scratch, w.fb_tickets, d_idx, d_score, m->num_sms, ts)))
    return rc;
  if (timeline && pipelined) AVL_CUDA(cudaEventRecord(w.tl[par][5], ts));
  if (g_profiling && !pipelined) AVL_CUDA(cudaEventRecord(w.ev[3], s));
  if (!(flags & AVL_ON_DEVICE)) {
    AVL_CUDA(cudaMemcpyAsync(out_idx, d_idx, sizeof(int64_t) * nq * k, cudaMemcpyDeviceToHost, ts));
    AVL_CUDA(cudaMemcpyAsync(out_score, d_score, sizeof(float) * nq * k, cudaMemcpyDeviceToHost, ts));
  }
  if (pipelined) {
    AVL_CUDA(cudaEventRecord(w.ev_tail[par], ts));
    w.tail_pending[par] = true;
    ++w.n_pipelined;
    // the PREVIOUS call's results become ordered on `stream` here, behind this call's screen; this call's own results
    // follow with the next call or with avl_map_flush
    if (w.tail_pending[par ^ 1]) AVL_CUDA(cudaStreamWaitEvent(s, w.ev_tail[par ^ 1], 0));
    if (!(flags & AVL_ON_DEVICE) && !(flags & AVL_ASYNC)) {
      cudaError_t e = cudaStreamSynchronize(ts);
      if (e != cudaSuccess) return check_watchdog(m, cuda_fail(e, "topk (pipelined)", __FILE__, __LINE__));
    }
    return AVL_OK;
  }
  const bool triage = m->n > 0 && (p.debug_flags & 64);
  if (stats || triage) AVL_CUDA(cudaMemcpyAsync(w.pin, cand_cnt, sizeof(uint32_t) * 2 * AVL_MAX_QUERIES, cudaMemcpyDeviceToHost, s));
  // device-pointer calls without stats return here, asynchronously: results are ordered on `stream` like any kernel's
  if ((!(flags & AVL_ON_DEVICE) && !(flags & AVL_ASYNC)) || stats || triage) {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return check_watchdog(m, cuda_fail(e, "topk", __FILE__, __LINE__));
  }
  if (triage) {  // true SM clock of the main screen launch (cycles / nanoseconds of block 0) and per-CTA spans
    fprintf(stderr, "[avl clock] screen: %llu cycles in %u ns = %.0f MHz\n",
            static_cast<unsigned long long>(w.dbg_host[8]) | (static_cast<unsigned long long>(w.dbg_host[9]) << 32),
            w.dbg_host[10], w.dbg_host[10] ? 1e3 * (static_cast<double>(w.dbg_host[8]) + 4294967296.0 * w.dbg_host[9]) / w.dbg_host[10] : 0.0);
    uint32_t t0 = 0xFFFFFFFFu;
    for (int b = 0; b < grid && b < 160; ++b) t0 = std::min(t0, w.dbg_host[16 + 4 * b]);
    for (int b = 0; b < grid && b < 160; ++b) {
      const uint32_t* r = w.dbg_host + 16 + 4 * b;
      fprintf(stderr, "[avl cta] %3d sm %3u start %6u ns dur %7u ns end %7u ns %5.0f MHz tiles %u\n", b, r[3], r[0] - t0, r[1],
              r[0] - t0 + r[1], r[1] ? 1e3 * r[2] / r[1] : 0.0, w.dbg_host[16 + 4 * 160 + b]);
    }
  }
  if (stats) {
    const uint32_t* cnt = w.pin;
    const uint32_t* ovf = w.pin + AVL_MAX_QUERIES;
    int n_fallback = 0;
    int64_t n_cand = 0;
    for (int q = 0; q < nq; ++q) {
      n_cand += cnt[q];
      n_fallback += ovf[q] ? 1 : 0;
    }
    stats->cta_group = qs.cg;
    stats->n_launches = 6;  // query_prepare, sample screen, select, screen, finalize, fallback
    stats->n_candidates = n_cand;
    stats->n_fallback_queries = n_fallback;
    stats->sample_rows = static_cast<int32_t>(std::min<int64_t>(n_sample, m->n));
    if (g_profiling && m->n > 0) {
      cudaEventElapsedTime(&stats->ms_screen, w.ev[1], w.ev[2]);
      cudaEventElapsedTime(&stats->ms_total, w.ev[0], w.ev[3]);
    }
  }
  return AVL_OK;
}

int avl_map_flush(avl_map* m, void* stream) {
  AVL_ARG(m != nullptr, "map is NULL");
  return drain_tails(m, static_cast<cudaStream_t>(stream));
}

int avl_map_tail_stream(avl_map* m, void** out_stream) {
  AVL_ARG(m != nullptr && out_stream != nullptr, "NULL argument");
  *out_stream = m->ws.tail_stream;
  return AVL_OK;
}

int avl_map_screen_times(avl_map* m, float* out_ms, int32_t cap, int32_t* n_out) {
  AVL_ARG(m != nullptr && out_ms != nullptr && n_out != nullptr && cap >= 0, "invalid argument");
  Workspace& w = m->ws;
  const int have = std::min(w.ring_n, 256);
  const int n = std::min<int>(have, cap);
  *n_out = 0;
  if (have > 0) AVL_CUDA(cudaEventSynchronize(w.ring_ev[2 * ((w.ring_n - 1) % 256) + 1]));
  for (int i = 0; i < n; ++i) {
    const int slot = (w.ring_n - n + i) % 256;
    AVL_CUDA(cudaEventElapsedTime(out_ms + i, w.ring_ev[2 * slot], w.ring_ev[2 * slot + 1]));
  }
  *n_out = n;
  w.ring_n = 0;
  return AVL_OK;
}

int avl_merge_topk(const int64_t* idx, const float* val, int32_t n_shards, int32_t nq, int32_t k,
                   int64_t* out_idx, float* out_val, int flags, void* stream) {
  AVL_ARG(idx && val && out_idx && out_val, "NULL argument");
  AVL_ARG(n_shards >= 1 && nq >= 1 && k >= 1 && k <= AVL_MAX_TOPK, "invalid shape");
  if (!(flags & AVL_ON_DEVICE)) {
    set_error("avl_merge_topk takes device pointers (the gathered NCCL buffer)");
    return AVL_ERR_UNSUPPORTED;
  }
  return launch_merge_topk(idx, val, n_shards, nq, k, out_idx, out_val, static_cast<cudaStream_t>(stream));
}

// exact path: fp64-accumulated dense columns of both modalities, min-max, combine, vector top-k per pair
static int fuse_topk_exact(avl_map* ma, const float* qa, const float* scale_a, int normalize_a, avl_map* mb,
                           const float* qb, const float* scale_b, int normalize_b, int32_t n_pairs, int32_t combine,
                           int32_t k, int64_t* out_idx, float* out_heat, int flags, void* stream) {
  AVL_ARG(ma && mb && qa && qb && out_idx && out_heat, "NULL argument");
  AVL_ARG(ma->n == mb->n, "both maps must have the same number of rows");
  AVL_ARG(n_pairs >= 1 && n_pairs <= AVL_MAX_QUERIES, "n_pairs out of range");
  AVL_ARG(k >= 1 && k <= AVL_MAX_TOPK, "k must be in [1, AVL_MAX_TOPK]");
  AVL_ARG(combine >= AVL_FUSE_PRODUCT && combine <= AVL_FUSE_SUM, "unknown combine rule");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = ma->n;
  QuerySetup sa, sb;
  int rc;
  if ((rc = setup_queries(ma, qa, n_pairs, scale_a, flags, 0, -1, s, &sa))) return rc;
  if ((rc = setup_queries(mb, qb, n_pairs, scale_b, flags, 0, -1, s, &sb))) return rc;
  if ((rc = ensure_column(ma))) return rc;
  float *da = nullptr, *db = nullptr, *mm = nullptr;
  const size_t elems = static_cast<size_t>(std::max<int64_t>(n, 1)) * n_pairs;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&da), elems * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&db), elems * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&mm), 4 * n_pairs * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(da); cudaFree(db); cudaFree(mm);
    return cuda_fail(e, "fuse buffers", __FILE__, __LINE__);
  }
  int64_t* d_idx = (flags & AVL_ON_DEVICE) ? out_idx : ma->ws.out_idx;
  float* d_heat = (flags & AVL_ON_DEVICE) ? out_heat : ma->ws.out_score;
  do {
    // column-major (pair, row) exact scores of both modalities
    if ((rc = launch_dense_exact(ma->feat, n, ma->d, sa.q_dev, n_pairs, sa.scale_dev, ma->row_norm, normalize_a,
                                 da, 1, n, s))) break;
    if ((rc = launch_dense_exact(mb->feat, n, mb->d, sb.q_dev, n_pairs, sb.scale_dev, mb->row_norm, normalize_b,
                                 db, 1, n, s))) break;
    if ((rc = launch_minmax_cols(da, n, n_pairs, mm, mm + n_pairs, s))) break;
    if ((rc = launch_minmax_cols(db, n, n_pairs, mm + 2 * n_pairs, mm + 3 * n_pairs, s))) break;
    for (int j = 0; j < n_pairs && rc == AVL_OK; ++j) {
      rc = launch_fuse_heat(da, db, n, j, n_pairs, mm, mm + n_pairs, mm + 2 * n_pairs, mm + 3 * n_pairs, combine,
                            ma->ws.column, s);
      if (rc) break;
      rc = launch_topk_vector(ma->ws.column, n, k, d_idx + static_cast<size_t>(j) * k,
                              d_heat + static_cast<size_t>(j) * k, ma->ws.topk_scratch, ma->ws.topk_scratch_bytes, s);
    }
    if (rc) break;
    if (!(flags & AVL_ON_DEVICE)) {
      e = cudaMemcpyAsync(out_idx, d_idx, sizeof(int64_t) * n_pairs * k, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(out_heat, d_heat, sizeof(float) * n_pairs * k, cudaMemcpyDeviceToHost, s);
      if (e != cudaSuccess) { rc = cuda_fail(e, "fuse D2H", __FILE__, __LINE__); break; }
    }
  } while (0);
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess && rc == AVL_OK) rc = cuda_fail(e, "fuse", __FILE__, __LINE__);
  cudaFree(da); cudaFree(db); cudaFree(mm);
  return rc;
}

// screened path: two dense tcgen05 passes + interval propagation + exact re-score of the survivors (sim_exact.cu,
// "cross-modal fusion through the screen").  Returns 1 when the call must be redone by the exact path
// (candidate-list overflow: massive ties, degenerate columns).
static int fuse_topk_screened(avl_map* ma, const float* qa, const float* scale_a, int normalize_a, avl_map* mb,
                              const float* qb, const float* scale_b, int normalize_b, int32_t n_pairs, int32_t combine,
                              int32_t k, int64_t* out_idx, float* out_heat, int flags, cudaStream_t s) {
  const int64_t n = ma->n;
  QuerySetup sa, sb;
  int rc;
  if ((rc = setup_queries(ma, qa, n_pairs, scale_a, flags, 0, kModeDense, s, &sa))) return rc;
  if ((rc = setup_queries(mb, qb, n_pairs, scale_b, flags, 0, kModeDense, s, &sb))) return rc;
  Workspace& w = ma->ws;
  const uint32_t cand_cap = 8192, ext_cap = 2048;
  if ((rc = ensure_cands(ma, cand_cap))) return rc;
  // sample for the heat threshold: same sizing rule as avl_sim_topk
  int64_t n0 = std::max<int64_t>(n / 64, static_cast<int64_t>(k) * n / 1024);
  n0 = std::min<int64_t>(std::max<int64_t>(n0, 8192), n);
  const int64_t stride = std::max<int64_t>(1, n / n0);
  const int32_t n_sample = static_cast<int32_t>((n + stride - 1) / stride);
  if ((rc = ensure_sample(ma, static_cast<size_t>(n_sample) * n_pairs))) return rc;
  const size_t elems = static_cast<size_t>(n) * n_pairs;
  if (w.fuse_elems < elems) {
    cudaFree(w.fuse_a); cudaFree(w.fuse_b);
    w.fuse_a = w.fuse_b = nullptr;
    w.fuse_elems = 0;
    if ((rc = dev_alloc(&w.fuse_a, elems, &ma->bytes))) return rc;
    if ((rc = dev_alloc(&w.fuse_b, elems, &ma->bytes))) return rc;
    w.fuse_elems = elems;
  }
  if (!w.fuse_small) {
    // max_lb[512] min_ub[512] ext_cnt[1024] mm[1024 floats] overflow[1] | ext_row[1024][ext_cap]
    if ((rc = dev_alloc(&w.fuse_small, static_cast<size_t>(3200) + static_cast<size_t>(4) * AVL_MAX_QUERIES * ext_cap, &ma->bytes)))
      return rc;
  }
  for (int side = 0; side < 2; ++side) {
    avl_map* m = side ? mb : ma;
    const QuerySetup& qs = side ? sb : sa;
    ScreenParams p;
    base_params(m, qs, n_pairs, 0, &p);
    p.mode = kModeDense;
    p.dense_out = side ? w.fuse_b : w.fuse_a;
    p.dense_rs = 1;
    p.dense_cs = n;
    p.dense_cols = n_pairs;
    if ((rc = run_screen(m, qs, p, s))) return rc;
  }
  FuseSideHost ha{w.fuse_a, ma->row_c, ma->row_an, ma->row_norm, ma->ws.q_bn, ma->ws.q_glob, sa.scale_dev, ma->feat, sa.q_dev, ma->d, normalize_a};
  FuseSideHost hb{w.fuse_b, mb->row_c, mb->row_an, mb->row_norm, mb->ws.q_bn, mb->ws.q_glob, sb.scale_dev, mb->feat, sb.q_dev, mb->d, normalize_b};
  FuseScratch fs;
  fs.max_lb = w.fuse_small;
  fs.min_ub = w.fuse_small + 512;
  fs.ext_cnt = w.fuse_small + 1024;
  fs.mm = reinterpret_cast<float*>(w.fuse_small + 2048);
  fs.overflow = w.fuse_small + 3072;
  fs.ext_row = w.fuse_small + 3200;
  fs.ext_cap = ext_cap;
  fs.sample_t = w.sample_t;
  fs.n_sample = n_sample;
  fs.sample_stride = stride;
  fs.thr = w.thr_t;
  fs.cand_cnt = w.cand_cnt;
  fs.cand_row = w.cand_row;
  fs.cand_cap = cand_cap;
  fs.cand_hi = w.cand_val;
  if (!w.cand_val2 && (rc = dev_alloc(&w.cand_val2, static_cast<size_t>(AVL_MAX_QUERIES) * w.cand_cap, &ma->bytes))) return rc;
  fs.cand_lo = w.cand_val2;
  int64_t* d_idx = (flags & AVL_ON_DEVICE) ? out_idx : w.out_idx;
  float* d_heat = (flags & AVL_ON_DEVICE) ? out_heat : w.out_score;
  if ((rc = launch_fuse_screened(ha, hb, n, n_pairs, combine, k, fs, d_idx, d_heat, ma->num_sms, s))) return rc;
  uint32_t ovf = 0;
  AVL_CUDA(cudaMemcpyAsync(&ovf, fs.overflow, sizeof(ovf), cudaMemcpyDeviceToHost, s));
  if (!(flags & AVL_ON_DEVICE)) {
    AVL_CUDA(cudaMemcpyAsync(out_idx, d_idx, sizeof(int64_t) * n_pairs * k, cudaMemcpyDeviceToHost, s));
    AVL_CUDA(cudaMemcpyAsync(out_heat, d_heat, sizeof(float) * n_pairs * k, cudaMemcpyDeviceToHost, s));
  }
  {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return check_watchdog(ma, cuda_fail(e, "fuse (screened)", __FILE__, __LINE__));
  }
  if (getenv("AVL_FUSE_DEBUG")) {
    std::vector<uint32_t> cnt(n_pairs), ext(4 * n_pairs);
    cudaMemcpy(cnt.data(), fs.cand_cnt, sizeof(uint32_t) * n_pairs, cudaMemcpyDeviceToHost);
    cudaMemcpy(ext.data(), fs.ext_cnt, sizeof(uint32_t) * 4 * n_pairs, cudaMemcpyDeviceToHost);
    uint64_t tc = 0, te = 0;
    uint32_t mc = 0, me = 0;
    for (uint32_t c : cnt) { tc += c; mc = std::max(mc, c); }
    for (uint32_t c : ext) { te += c; me = std::max(me, c); }
    fprintf(stderr, "[avl fuse] heat candidates: mean %.0f max %u per pair (cap %u); extreme candidates: mean %.1f max %u (cap %u); overflow %u\n",
            static_cast<double>(tc) / n_pairs, mc, cand_cap, static_cast<double>(te) / (4 * n_pairs), me, ext_cap, ovf);
  }
  return ovf ? 1 : AVL_OK;
}

int avl_fuse_topk(avl_map* ma, const float* qa, const float* scale_a, int normalize_a, avl_map* mb,
                  const float* qb, const float* scale_b, int normalize_b, int32_t n_pairs, int32_t combine,
                  int32_t k, int64_t* out_idx, float* out_heat, int flags, void* stream) {
  AVL_ARG(ma && mb && qa && qb && out_idx && out_heat, "NULL argument");
  AVL_ARG(ma->n == mb->n, "both maps must have the same number of rows");
  AVL_ARG(n_pairs >= 1 && n_pairs <= AVL_MAX_QUERIES, "n_pairs out of range");
  AVL_ARG(k >= 1 && k <= AVL_MAX_TOPK, "k must be in [1, AVL_MAX_TOPK]");
  AVL_ARG(combine >= AVL_FUSE_PRODUCT && combine <= AVL_FUSE_SUM, "unknown combine rule");
  int rc = scale_positive(scale_a, n_pairs, flags);
  if (rc == AVL_OK) rc = scale_positive(scale_b, n_pairs, flags);
  if (rc == AVL_OK) rc = drain_tails(ma, static_cast<cudaStream_t>(stream));
  if (rc == AVL_OK) rc = drain_tails(mb, static_cast<cudaStream_t>(stream));
  const bool force_exact = getenv("AVL_FUSE_EXACT") != nullptr;  // A/B and tests of the exact path
  const bool can_screen = rc == AVL_OK && !force_exact && ma != mb && ma->n >= 1024 &&
                          ma->n < (int64_t(1) << 32) - 1;
  if (can_screen) {
    rc = fuse_topk_screened(ma, qa, scale_a, normalize_a, mb, qb, scale_b, normalize_b, n_pairs, combine, k, out_idx,
                            out_heat, flags, static_cast<cudaStream_t>(stream));
    if (rc != 1) return rc;  // done, or a real error
  }
  return fuse_topk_exact(ma, qa, scale_a, normalize_a, mb, qb, scale_b, normalize_b, n_pairs, combine, k, out_idx,
                         out_heat, flags, stream);
}

}  // extern "C"
