// Internal declarations shared by the translation units of libavlmaps_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>

#include "../../include/avlmaps_b200.h"

namespace avl {

// ---- error plumbing ---------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define AVL_CUDA(expr)                                                      \
  do {                                                                      \
    cudaError_t _e = (expr);                                                \
    if (_e != cudaSuccess) return ::avl::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define AVL_ARG(cond, msg)               \
  do {                                   \
    if (!(cond)) {                       \
      ::avl::set_error(msg);             \
      return AVL_ERR_ARG;                \
    }                                    \
  } while (0)

// ---- programmatic dependent launch ------------------------------------------
// The kernels of one top-k call (query_prepare, sample screen, select, screen, finalize, fallback) are launched with
// programmatic stream serialization: a kernel's CTAs may be scheduled while its predecessor drains, run their
// data-independent prologue, and block in pdl_wait() until the predecessor has completed and flushed -- the launch
// latency and the prologue leave the critical path (AVL_PDL=0 launches them the ordinary way).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- screen kernel (sim_screen.cu) -----------------------------------------
enum ScreenMode : int32_t { kModeDense = 0, kModeArgmax = 1, kModeThresh = 2 };

constexpr int kTileRows = 128;   // rows (voxels) per CTA per tile = UMMA M per CTA
constexpr int kBlockK = 64;      // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kStageBytes = kTileRows * kBlockK * 2;  // 16 KiB of A per pipeline stage
constexpr int kFlagWords = 8;    // candidate bitmask words per flagged row (256 queries)

struct ScreenParams {
  int64_t n_rows;        // rows of the map
  int32_t kblocks;       // dpad / 64
  int32_t nq;            // valid query columns
  int32_t npad;          // UMMA N (multiple of 16, <= 256); B has npad rows
  int32_t num_tiles;     // row tiles (of 128*CG rows) this launch processes
  int32_t tile_stride;   // launch tile j covers map tile j * tile_stride (sampling)
  int32_t stages;        // A pipeline depth
  int32_t mode;          // ScreenMode
  int32_t normalize;     // divide by the fp32 row norm
  uint32_t* tile_ctr;     // global tile counter of the dynamic schedule (never reset)
  uint32_t tile_base;     // its value before this launch's first fetch
  int32_t a_tiled;        // operand copy of the map is tile-major ([tile][k-block][128 rows][64]), see map_prepare
  int32_t debug_flags;    // perf triage only (AVL_DEBUG_FLAGS): 1 = no MMA, 2 = no A loads, 4 = no epilogue work,
                          // 8 = pre-test always passes, 16 = thresholds +inf (nothing emitted), 32 = TMEM loads only
  int32_t op_f16;         // operands are fp16 instead of bf16 (same tcgen05 kind::f16 rate, 8x smaller rounding residual)
  // per-row statistics (map_prepare): see DESIGN.md "error band"
  const float* row_norm;   // ||a_i||  (fp32 row, fp64-accumulated)
  const float* row_c;      // >= ||a_i - bf16(a_i)|| + kappa * ||bf16(a_i)||
  const float* row_an;     // >= ||bf16(a_i)||
  // per-query statistics (query_prepare)
  const float* q_bn;       // >= ||b_q||
  const __nv_bfloat16* a_base; // operand copy of the map (tile-major: the L2 prefetch addresses it directly)
  const __nv_bfloat16* bq; // bf16 queries (256 zero-padded rows x dpad)
  const float* q_glob;     // [0] rho >= max_q ||b_q - bf16(b_q)|| / ||b_q|| (+slack), [1] max_q q_bn
  // kModeDense
  float* dense_out;        // element (r, q) at r * dense_rs + q * dense_cs, r = compact row
  int64_t dense_rs, dense_cs;
  int32_t dense_cols;      // columns to store (nq or npad)
  int32_t dense_lb;        // store lower bounds (s~ - eps)/w instead of s~ (rows past n_rows: -inf)
  // kModeArgmax
  int32_t* argmax_out;     // (n_rows,)
  uint32_t* flag_count;    // [1]
  uint32_t* flag_rows;     // [flag_cap]
  uint32_t* flag_masks;    // [flag_cap][kFlagWords]
  uint32_t flag_cap;
  // kModeThresh
  const float* thr_t;      // per query tau_q / scale_q  (+inf disables a query)
  uint32_t* cand_cnt;      // [grid][AVL_MAX_QUERIES] fill count of bucket (query, CTA); may exceed cand_bucket
  uint32_t* cand_row;      // [nq][grid][cand_bucket] map row
  float* cand_val;         // [nq][grid][cand_bucket] screen score s~
  uint32_t cand_bucket;    // entries per (query, CTA) bucket
  // watchdog record (host-mapped), may be null
  uint32_t* dbg;
};

// Launch the tcgen05 screen.  tmap_a / tmap_b are CUtensorMap (128 bytes each).
int launch_screen(int cta_group, const void* tmap_a, const void* tmap_b, const ScreenParams& p,
                  int num_sms, size_t smem_bytes, cudaStream_t stream);
size_t screen_smem_bytes(int cta_group, int npad, int kblocks, int stages, int mode);
int screen_pick_stages(int cta_group, int npad, int kblocks, int mode);  // <=0: does not fit
int screen_grid(int cta_group, int num_sms, int num_tiles);  // CTAs launch_screen starts

// ---- exact / helper kernels (sim_exact.cu) --------------------------------
int launch_map_prepare(const float* feat, int64_t n, int32_t d, int32_t dpad, __nv_bfloat16* bf,
                       float* row_norm, float* row_c, float* row_an, float kappa, int f16, uint32_t* nonfinite,
                       int tiled, cudaStream_t s);
int launch_query_prepare(const float* q, const float* fold_scale, int32_t nq, int32_t d, int32_t dpad,
                         int32_t npad, __nv_bfloat16* bq, float* q_bn, float* q_glob, int f16, cudaStream_t s);
int launch_dense_exact(const float* feat, int64_t n, int32_t d, const float* q, int32_t nq,
                       const float* scale, const float* row_norm, int normalize, float* out,
                       int64_t out_rs, int64_t out_cs, cudaStream_t s);
int launch_column_exact(const float* feat, int64_t n, int32_t d, const float* q, const float* scale,
                        const float* row_norm, int normalize, float* out, int num_sms, cudaStream_t s);
int launch_argmax_rerank(const float* feat, int32_t d, const float* q, double* q64, const float* q_bn_raw, int32_t nq,
                         const float* scale,
                         const float* row_norm, int normalize, const uint32_t* flag_count,
                         const uint32_t* flag_rows, const uint32_t* flag_masks, uint32_t flag_cap,
                         int32_t* argmax_out, int num_sms, cudaStream_t s);
int launch_select_threshold(const float* sample_lb, int32_t n_sample_rows, int64_t ld, int32_t nq, int32_t k,
                            float* thr_t, cudaStream_t s);
int launch_topk_finalize(const float* feat, int64_t n_rows, int32_t d, const float* q, int32_t nq,
                         const float* scale, const float* row_norm, const float* row_c,
                         const float* row_an, const float* q_bn, const float* q_glob, int normalize,
                         int32_t k, const uint32_t* bucket_cnt, int32_t grid, uint32_t cand_bucket,
                         const uint32_t* cand_row, const float* cand_val, uint32_t fin_cap, int64_t* out_idx,
                         float* out_score, uint32_t* cand_total, uint32_t* overflow_flags, uint32_t* gscratch,
                         cudaStream_t s);
// exact re-score of the queries whose overflow flag is set, decided on the device (no host round trip)
int launch_topk_fallback(const float* feat, int64_t n, int32_t d, const float* q, int32_t nq, const float* scale,
                         const float* row_norm, int normalize, int32_t k, const uint32_t* overflow_flags,
                         void* scratch, uint32_t* tickets, int64_t* out_idx, float* out_score, int num_sms,
                         cudaStream_t s);
size_t topk_fallback_scratch_bytes(int num_sms);
int launch_topk_vector(const float* values, int64_t n, int32_t k, int64_t* out_idx, float* out_val,
                       void* scratch, size_t scratch_bytes, cudaStream_t s);
size_t topk_vector_scratch_bytes(int64_t n);
size_t topk_finalize_smem(uint32_t cand_cap);
int launch_merge_topk(const int64_t* idx, const float* val, int32_t n_shards, int32_t nq, int32_t k,
                      int64_t* out_idx, float* out_val, cudaStream_t s);
int launch_minmax_cols(const float* m, int64_t n, int32_t cols, float* out_min, float* out_max,
                       cudaStream_t s);
int launch_fuse_heat(const float* sa, const float* sb, int64_t n, int32_t pair, int32_t cols,
                     const float* min_a, const float* max_a, const float* min_b, const float* max_b,
                     int32_t combine, float* heat, cudaStream_t s);


// ---- cross-modal fusion through the screen (sim_exact.cu)
struct FuseSideHost {
  const float* dense;     // (pairs, n) column-major screen scores of this modality
  const float* row_c; const float* row_an; const float* row_norm;
  const float* q_bn; const float* q_glob; const float* scale;
  const float* feat; const float* q;
  int32_t d; int32_t normalize;
};
struct FuseScratch {
  uint32_t* max_lb; uint32_t* min_ub;      // [2 * pairs] ordered-uint column statistics
  uint32_t* ext_cnt; uint32_t* ext_row;    // [4 * pairs], [4 * pairs][ext_cap] candidates of the column max / min
  uint32_t ext_cap;
  float* mm;                               // [4][pairs]: min a, max a, min b, max b (exact)
  float* sample_t; int32_t n_sample; int64_t sample_stride;
  float* thr;                              // [pairs]
  uint32_t* cand_cnt; uint32_t* cand_row; uint32_t cand_cap;
  float* cand_lo; float* cand_hi;          // [pairs][cand_cap] heat bounds of the candidates
  uint32_t* overflow;                      // [1]
};
int launch_fuse_screened(const FuseSideHost& a, const FuseSideHost& b, int64_t n, int32_t pairs, int32_t combine,
                         int32_t k, const FuseScratch& w, int64_t* out_idx, float* out_heat, int num_sms,
                         cudaStream_t s);

}  // namespace avl
