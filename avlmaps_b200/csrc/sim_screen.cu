// Voxel x query similarity screen: S~ = bf16(A) . bf16(B)^T on the 5th-gen tensor cores.
//
//   A  = voxel features  (n_rows x dpad bf16, K-major)   streamed once from HBM by TMA
//   B  = query embeddings (npad   x dpad bf16, K-major)   resident in shared memory
//   S~ = 128-row x npad accumulator tiles in TMEM, consumed in place by a fused epilogue
//
// Replaces the OpenBLAS sgemm behind `map_feats @ text_feats.T`
// (reference avlmaps/utils/clip_utils.py:227-229) plus the argmax / selection that follows it
// (avlmaps/map/vlmap.py:123-124, avlmaps/robot/habitat_lang_robot.py:427-430); the (N, Q) score
// matrix never exists in HBM.
//
// Warp roles (256 threads): 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator, 4-7 = epilogue
// (warp w owns TMEM lanes 32*(w%4)..+31, i.e. one voxel row per thread).
// CG = 2 pairs two SMs (tcgen05 cta_group::2): UMMA M = 256, each CTA stages its own 128 voxel
// rows and HALF of B, so a 256-query x 512-d B (256 KiB) is resident across the pair.
#include <cuda.h>

#include <cstdlib>

#include "avl_internal.h"
#include "ptx_sm100.cuh"

namespace avl {

namespace {

constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter
constexpr int kThreads = 128 + kEpiWarps * 32;     // 384
constexpr int kMaxStages = 12;
constexpr int kCtrlBytes = 1024;   // barriers + tmem slot
constexpr int kQConstBytes = 2048; // float2[256]
constexpr int kRingEntries = 64;   // per epilogue warp: staged candidates before a 32-entry flush
constexpr int kRingBytes = kEpiWarps * kRingEntries * 12;
constexpr int kXchgBytes = 4 * 32 * 8 * 4;  // argmax: the two warps of a lane quarter exchange keys / mask words
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccumStride = 256;  // columns between the two accumulator stages

__device__ __forceinline__ uint32_t f2ord(uint32_t b) {
  // monotone map float bits -> uint32 (larger float <=> larger uint)
  return b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
  return __uint_as_float(b);
}

struct TileCtx {
  int64_t row;    // map row of this thread
  int64_t crow;   // compact row (dense output of sampled launches)
  bool valid;
  uint32_t taddr; // TMEM address of column 0 of this thread's row
};

// ---- epilogue: dense store ------------------------------------------------------------
// dense_lb: store the LOWER bound of the exact score in threshold units, (s~ - r_i ||b_q||) / w_i
// (directed rounding), instead of s~ -- what the sampled threshold pass needs.
template <int W>
__device__ __forceinline__ void dense_chunk(const ScreenParams& p, const TileCtx& t, int c0, const float2* qc,
                                            float r_i, float w_i) {
  uint32_t v[W];
  if constexpr (W == 32) ptx::tmem_ld32(t.taddr + c0, v); else ptx::tmem_ld16(t.taddr + c0, v);
  ptx::tmem_ld_wait();
  float* o = p.dense_out + t.crow * p.dense_rs + static_cast<int64_t>(c0) * p.dense_cs;
  if (p.dense_lb == 2) {
    // threshold sample, reduced in place: ONE value per 32-row group (this warp's rows) and query, the max of the
    // lower bounds.  The selection only ever uses maxima of disjoint row groups (select_threshold_kernel), so the
    // (rows x queries) sample matrix never has to exist: 32x less to write here and to read there.
    const int lane = threadIdx.x & 31;
    uint32_t mine = 0u;  // lane j keeps column c0 + j; 0 = no valid row in the group
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const float x = (__uint_as_float(v[j]) - r_i * qc[c0 + j].y) * w_i;
      const float lb = fmaf(-fabsf(x), 9.5367431640625e-7f, x);
      const uint32_t key = t.valid ? max(f2ord(__float_as_uint(lb)), 1u) : 0u;
      const uint32_t m = __reduce_max_sync(0xffffffffu, key);
      if (lane == j) mine = m;
    }
    if (lane < W && c0 + lane < p.dense_cols)
      p.dense_out[(t.crow >> 5) * p.dense_rs + static_cast<int64_t>(c0 + lane) * p.dense_cs] =
          mine ? ord2f(mine) : -INFINITY;
  } else if (p.dense_lb) {
    // rows past the end of the map store -inf so that they never raise a threshold
#pragma unroll
    for (int j = 0; j < W; ++j) {
      if (c0 + j < p.dense_cols) {
        // (s~ - r ||b||) / w with the rounding pushed DOWN by a relative 2^-20 instead of directed-rounding
        // divides (the threshold only has to be a lower bound)
        const float x = (__uint_as_float(v[j]) - r_i * qc[c0 + j].y) * w_i;  // w_i holds 1 / w here
        o[j * p.dense_cs] = t.valid ? fmaf(-fabsf(x), 9.5367431640625e-7f, x) : -INFINITY;
      }
    }
  } else if (t.valid) {
#pragma unroll
    for (int j = 0; j < W; ++j)
      if (c0 + j < p.dense_cols) o[j * p.dense_cs] = __uint_as_float(v[j]);
  }
}

// ---- epilogue: per-row top-2 keys -------------------------------------------------------
template <int W>
__device__ __forceinline__ void argmax_chunk(const ScreenParams& p, const TileCtx& t, int c0,
                                             uint32_t& best, uint32_t& second) {
  uint32_t v[W];
  if constexpr (W == 32) ptx::tmem_ld32(t.taddr + c0, v); else ptx::tmem_ld16(t.taddr + c0, v);
  ptx::tmem_ld_wait();
  if (c0 + W <= p.nq) {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      uint32_t key = (f2ord(v[j]) & 0xFFFFFF00u) | static_cast<uint32_t>(255 - (c0 + j));
      second = max(second, min(best, key));
      best = max(best, key);
    }
  } else {
#pragma unroll
    for (int j = 0; j < W; ++j) {
      if (c0 + j < p.nq) {
        uint32_t key = (f2ord(v[j]) & 0xFFFFFF00u) | static_cast<uint32_t>(255 - (c0 + j));
        second = max(second, min(best, key));
        best = max(best, key);
      }
    }
  }
}

template <int W>
__device__ __forceinline__ uint32_t mask_chunk(const TileCtx& t, int c0, float thr) {
  uint32_t v[W];
  if constexpr (W == 32) ptx::tmem_ld32(t.taddr + c0, v); else ptx::tmem_ld16(t.taddr + c0, v);
  ptx::tmem_ld_wait();
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < W; ++j) m |= (__uint_as_float(v[j]) >= thr ? 1u : 0u) << j;
  return m;
}

// ---- epilogue: threshold screen ---------------------------------------------------------
// candidate (row i, query q)  <=>  upper bound of the exact score reaches the threshold:
//     (s~ + r_i * ||b_q||) / w_i  >=  T_q          (w_i = ||a_i|| if normalize_map else 1)
// Fast path: one predicate bit per score, no branch.  qc[q] = (T_q, ||b_q||) in shared memory.
template <int W, bool kNorm>
__device__ __forceinline__ uint32_t thresh_mask_chunk(const TileCtx& t, int c0, const float2* qc, float iw,
                                                      float ri, const float2 qk) {
  uint32_t v[W];
  if constexpr (W == 32) ptx::tmem_ld32(t.taddr + c0, v); else ptx::tmem_ld16(t.taddr + c0, v);
  ptx::tmem_ld_wait();
  // Pre-test on the row's maximum over the chunk against the chunk's smallest threshold and largest ||b||
  // (qk): one FMNMX per score instead of FFMA + compare + shift/or.  Candidates are ~0.04 % of the scores, so
  // ~99 % of the (row, chunk) pairs stop here; exactly conservative (max, fma and the comparison are monotone).
  {
    float vmax = __uint_as_float(v[0]);
#pragma unroll
    for (int j = 1; j < W; ++j) vmax = fmaxf(vmax, __uint_as_float(v[j]));
    const float u = kNorm ? fmaf(ri, qk.y, vmax * iw) : fmaf(ri, qk.y, vmax);
    if (!(u >= qk.x)) return 0u;  // also taken by rows past the map (ri = NaN)
  }
  const float4* qc4 = reinterpret_cast<const float4*>(qc + c0);
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < W; j += 2) {
    const float4 c = qc4[j >> 1];  // two queries per 16-byte broadcast load
    float u0, u1;
    if constexpr (kNorm) {
      u0 = fmaf(ri, c.y, __uint_as_float(v[j]) * iw);       // ri = r_i / w_i, iw = 1 / w_i
      u1 = fmaf(ri, c.w, __uint_as_float(v[j + 1]) * iw);
    } else {
      u0 = fmaf(ri, c.y, __uint_as_float(v[j]));
      u1 = fmaf(ri, c.w, __uint_as_float(v[j + 1]));
    }
    m |= (u0 >= c.x ? 1u : 0u) << j;
    m |= (u1 >= c.z ? 1u : 0u) << (j + 1);
  }
  return m;
}

// Slow path (rare, ~3 scores per warp and tile).  Kept tiny on purpose: an unrolled per-bit version
// thrashed the instruction cache, and one global atomic per candidate cost ~2000 cycles each.
// Marked scores are staged in a per-warp shared-memory ring and written to ONE global list in
// 32-entry bursts (one atomicAdd per burst); the finalize kernel regroups them by query.
struct Ring {
  uint32_t* row;   // [kRingEntries]
  uint32_t* q;     // [kRingEntries]
  float* val;      // [kRingEntries]
};

__device__ __noinline__ uint32_t ring_flush(const ScreenParams& p, Ring r, uint32_t pend, uint32_t count,
                                            uint32_t lane) {
  // every lane appends one staged entry to its query's list: 32 independent atomics in flight, one
  // round trip per 32 candidates (a per-candidate atomic in the epilogue cost ~2000 cycles each)
  if (lane < count) {
    const uint32_t q = r.q[lane];
    const uint32_t slot = atomicAdd(p.cand_cnt + q, 1u);
    if (slot < p.cand_cap) {
      p.cand_row[static_cast<size_t>(q) * p.cand_cap + slot] = r.row[lane];
      p.cand_val[static_cast<size_t>(q) * p.cand_cap + slot] = r.val[lane];
    }
  }
  __syncwarp();
  const uint32_t rem = pend - count;
  uint32_t a = 0, b = 0;
  float c = 0.f;
  if (lane < rem) { a = r.row[count + lane]; b = r.q[count + lane]; c = r.val[count + lane]; }
  __syncwarp();
  if (lane < rem) { r.row[lane] = a; r.q[lane] = b; r.val[lane] = c; }
  __syncwarp();
  return rem;
}

__device__ __forceinline__ uint32_t thresh_emit_word(const ScreenParams& p, uint32_t taddr, int64_t row, int c0,
                                                     uint32_t m, Ring r, uint32_t pend, uint32_t lane) {
  uint32_t u = __reduce_or_sync(0xffffffffu, m);
  while (u) {  // warp-uniform loop over the columns any lane marked
    const int j = __ffs(u) - 1;
    u &= u - 1;
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr + c0 + j) : "memory");
    ptx::tmem_ld_wait();
    const bool mine = (m >> j) & 1u;
    const uint32_t b = __ballot_sync(0xffffffffu, mine);
    if (mine) {
      const uint32_t slot = pend + __popc(b & ((1u << lane) - 1u));
      r.row[slot] = static_cast<uint32_t>(row);
      r.q[slot] = static_cast<uint32_t>(c0 + j);
      r.val[slot] = __uint_as_float(v);
    }
    pend += __popc(b);
    __syncwarp();
    if (pend >= 32u) pend = ring_flush(p, r, pend, 32u, lane);
  }
  return pend;
}

// One epilogue warp handles the 32-column words cb = half, half + 2, ... of its 32 rows.
template <bool kNorm>
__device__ __forceinline__ uint32_t thresh_tile(const ScreenParams& p, const TileCtx& t, const float2* qc,
                                                const float2* qchunk, float iw, float ri, int half, Ring r,
                                                uint32_t pend, uint32_t lane) {
  uint32_t m[kFlagWords / 2];
  uint32_t any = 0;
#pragma unroll
  for (int i = 0; i < kFlagWords / 2; ++i) {
    m[i] = 0u;
    const int c0 = (2 * i + half) * 32;
    if (c0 + 32 <= p.npad) m[i] = thresh_mask_chunk<32, kNorm>(t, c0, qc, iw, ri, qchunk[c0 >> 5]);
    else if (c0 < p.npad) m[i] = thresh_mask_chunk<16, kNorm>(t, c0, qc, iw, ri, qchunk[c0 >> 5]);
    any |= m[i];
  }
  if (__any_sync(0xffffffffu, any != 0u)) {
#pragma unroll
    for (int i = 0; i < kFlagWords / 2; ++i) {
      const int c0 = (2 * i + half) * 32;
      if (c0 < p.npad) pend = thresh_emit_word(p, t.taddr, t.row, c0, m[i], r, pend, lane);
    }
  }
  return pend;
}

// SB ("streamed B", opt-in AVL_STREAM_B=1, not yet measured): B is not resident; its k-block travels with A's in every
// pipeline stage (re-read from L2, where EVICT_LAST keeps it), which frees the 128 KiB B took per CTA at 256 queries
// for a deeper ring -- more voxel bytes in flight per SM.  SB = false is the measured kernel, instruction for
// instruction (tools/sass_diff.py).
template <int CG, bool SB>
__global__ void __launch_bounds__(kThreads, 1)
screen_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
              const ScreenParams p) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment; the offset is identical in both CTAs of a pair.
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (CG == 2) rank = ptx::cluster_ctarank();
  const bool leader = (rank == 0);

  const uint32_t b_rows = static_cast<uint32_t>(p.npad) / CG;
  const uint32_t bblk = b_rows * 128u;                       // bytes of one resident k-block of B
  const uint32_t b_bytes = bblk * static_cast<uint32_t>(p.kblocks);
  const uint32_t stage_bytes = SB ? kStageBytes + bblk : static_cast<uint32_t>(kStageBytes);  // SB: A tile, then B k-block
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + (SB ? 0u : ((b_bytes + 1023u) & ~1023u));
  uint8_t* ctrl = smem_a + static_cast<uint32_t>(p.stages) * stage_bytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(ctrl);    // [kMaxStages]
  uint64_t* bar_empty = bar_full + kMaxStages;               // [kMaxStages]
  uint64_t* bar_tfull = bar_empty + kMaxStages;              // [2]
  uint64_t* bar_tempty = bar_tfull + 2;                      // [2]
  uint64_t* bar_bfull = bar_tempty + 2;                      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_bfull + 1);
  float2* qc = reinterpret_cast<float2*>(ctrl + kCtrlBytes);
  float2* qchunk = reinterpret_cast<float2*>(ctrl + 512);    // [8] per 32-query chunk: (min threshold, max ||b||)
  uint8_t* ring_base = ctrl + kCtrlBytes + kQConstBytes;
  uint32_t* xchg_base = reinterpret_cast<uint32_t*>(ring_base + kRingBytes);

  if constexpr (CG == 2) ptx::cluster_sync_all();  // both CTAs resident before the paired TMEM alloc

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(bar_full + i), CG);   // leader's expect_tx arrive (+ peer's arrive)
      ptx::mbar_init(ptx::smem_u32(bar_empty + i), 1);   // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(bar_tfull + i), 1);          // one tcgen05.commit
      ptx::mbar_init(ptx::smem_u32(bar_tempty + i), CG * kEpiWarps);  // one arrive per epilogue warp of the pair
    }
    ptx::mbar_init(ptx::smem_u32(bar_bfull), CG);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<CG>(ptx::smem_u32(tmem_slot), kTmemCols);
  // per-query constants of the threshold screen
  if (threadIdx.x < 256) {
    const int q = threadIdx.x;
    float2 c = make_float2(__int_as_float(0x7f800000), 0.f);  // +inf: padded column never passes
    if (q < p.nq) {
      c.x = (p.mode == kModeThresh && !(p.debug_flags & 3)) ? p.thr_t[q]
            : (p.mode == kModeThresh ? __int_as_float(0x7f800000) : 0.f);  // triage modes emit nothing
      c.y = p.q_bn[q];
    }
    qc[q] = c;
  } else if (threadIdx.x < 264) {
    // chunk summaries for the pre-test of the threshold epilogue, straight from global memory (qc is not visible yet)
    const int c0 = (static_cast<int>(threadIdx.x) - 256) * 32;
    float tmin = __int_as_float(0x7f800000), bmax = 0.f;
    if (p.mode == kModeThresh && !(p.debug_flags & 3)) {
      for (int q = c0; q < min(c0 + 32, p.nq); ++q) {
        tmin = fminf(tmin, p.thr_t[q]);
        bmax = fmaxf(bmax, p.q_bn[q]);
      }
    }
    if (p.debug_flags & 8) tmin = -INFINITY;  // A/B: pre-test always passes (AVL_DEBUG_FLAGS=8)
    qchunk[threadIdx.x - 256] = make_float2(tmin, bmax);
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_units = static_cast<int>(gridDim.x) / CG;
  const int unit = static_cast<int>(blockIdx.x) / CG;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if constexpr (!SB) {
        for (int kb = 0; kb < p.kblocks; ++kb)
          ptx::tma_load_2d<CG>(ptx::smem_u32(smem_b + kb * bblk), &tmap_b, ptx::smem_u32(bar_bfull),
                               kb * kBlockK, static_cast<int32_t>(rank * b_rows), ptx::kEvictLast);
        if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(bar_bfull), b_bytes * CG);
        else ptx::mbar_arrive_cluster(ptx::smem_u32(bar_bfull), 0);
      }

      uint32_t stage = 0, phase = 0;
      // The ring holds `stages` x 16 KiB per CTA, not enough bytes in flight to cover HBM latency at
      // full rate when B takes most of the shared memory; so the A tiles of the next
      // `prefetch_tiles` tiles are pulled into L2 ahead of the loads that will need them.
      for (int t = 0; t < p.prefetch_tiles; ++t) {
        const int jj = unit + t * num_units;
        if (jj < p.num_tiles) {
          const int64_t r0 = (static_cast<int64_t>(jj) * p.tile_stride * CG + rank) * kTileRows;
          for (int kb = 0; kb < p.kblocks; ++kb) ptx::tma_prefetch_2d(&tmap_a, kb * kBlockK, static_cast<int32_t>(r0));
        }
      }
      for (int j = unit; j < p.num_tiles; j += num_units) {
        const int64_t row0 = (static_cast<int64_t>(j) * p.tile_stride * CG + rank) * kTileRows;
        const int jp = j + p.prefetch_tiles * num_units;
        const int64_t rowp = (static_cast<int64_t>(jp) * p.tile_stride * CG + rank) * kTileRows;
        const bool do_pf = p.prefetch_tiles > 0 && jp < p.num_tiles;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          if (do_pf) ptx::tma_prefetch_2d(&tmap_a, kb * kBlockK, static_cast<int32_t>(rowp));
          ptx::mbar_wait(ptx::smem_u32(bar_empty + stage), phase ^ 1u, p.dbg, 0x10u + stage);
          if (!(p.debug_flags & 2)) {
            ptx::tma_load_2d<CG>(ptx::smem_u32(smem_a + stage * stage_bytes), &tmap_a,
                                 ptx::smem_u32(bar_full + stage), kb * kBlockK,
                                 static_cast<int32_t>(row0), ptx::kEvictFirst);
            if constexpr (SB)
              ptx::tma_load_2d<CG>(ptx::smem_u32(smem_a + stage * stage_bytes + kStageBytes), &tmap_b,
                                   ptx::smem_u32(bar_full + stage), kb * kBlockK,
                                   static_cast<int32_t>(rank * b_rows), ptx::kEvictLast);
            if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(bar_full + stage), stage_bytes * CG);
            else ptx::mbar_arrive_cluster(ptx::smem_u32(bar_full + stage), 0);
          } else {  // triage: barrier protocol only
            if (leader) ptx::mbar_arrive(ptx::smem_u32(bar_full + stage));
            else ptx::mbar_arrive_cluster(ptx::smem_u32(bar_full + stage), 0);
          }
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = p.op_f16 ? ptx::make_idesc_f16(kTileRows * CG, static_cast<uint32_t>(p.npad))
                                      : ptx::make_idesc_bf16(kTileRows * CG, static_cast<uint32_t>(p.npad));
      if constexpr (!SB) {
        ptx::mbar_wait(ptx::smem_u32(bar_bfull), 0, p.dbg, 0x20u);
        ptx::tc_fence_after();
      }
      uint32_t stage = 0, phase = 0, it = 0;
      for (int j = unit; j < p.num_tiles; j += num_units, ++it) {
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        ptx::mbar_wait(ptx::smem_u32(bar_tempty + as), aphase ^ 1u, p.dbg, 0x30u + as);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccumStride;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(bar_full + stage), phase, p.dbg, 0x40u + stage);
          ptx::tc_fence_after();
          const uint64_t a0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_a + stage * stage_bytes));
          const uint64_t b0 = ptx::make_kmajor_sw128_desc(
              ptx::smem_u32(SB ? smem_a + stage * stage_bytes + kStageBytes : smem_b + kb * bblk));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)  // UMMA K = 16 bf16 = 32 bytes inside the swizzle row
            if (!(p.debug_flags & 1))
              ptx::umma_bf16<CG>(tmem_d, a0 + 2u * k, b0 + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit<CG>(ptx::smem_u32(bar_empty + stage));   // frees the A stage in both CTAs
          if (kb == p.kblocks - 1) ptx::umma_commit<CG>(ptx::smem_u32(bar_tfull + as));
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (TMEM -> registers -> fused reduction) =====================
    const uint32_t lane_base = (warp & 3u) * 32u;
    const int half = static_cast<int>((warp - 4u) >> 2);  // which interleaved set of 32-column words
    const float rho = p.q_glob[0], bn_max = p.q_glob[1];
    Ring ring;
    {
      uint8_t* rb = ring_base + (warp - 4u) * (kRingEntries * 12);
      ring.row = reinterpret_cast<uint32_t*>(rb);
      ring.q = ring.row + kRingEntries;
      ring.val = reinterpret_cast<float*>(ring.q + kRingEntries);
    }
    uint32_t pend = 0;
    uint32_t it = 0;
    for (int j = unit; j < p.num_tiles; j += num_units, ++it) {
      const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
      TileCtx t;
      t.row = (static_cast<int64_t>(j) * p.tile_stride * CG + rank) * kTileRows + lane_base + lane;
      t.crow = (static_cast<int64_t>(j) * CG + rank) * kTileRows + lane_base + lane;
      t.valid = t.row < p.n_rows;
      t.taddr = tmem_base + as * kAccumStride + (lane_base << 16);
      const int n32 = p.npad & ~31;
      // this row's statistics: issued BEFORE waiting for the accumulator so that their global-memory
      // latency overlaps the wait instead of sitting between the MMAs and the release of the stage
      float st_an = 0.f, st_c = 0.f, st_norm = 1.f;
      if (t.valid) {
        st_an = p.row_an[t.row];
        st_c = p.row_c[t.row];
        if (p.normalize) st_norm = p.row_norm[t.row];
      }
      ptx::mbar_wait(ptx::smem_u32(bar_tfull + as), aphase, p.dbg, 0x50u + as);
      ptx::tc_fence_after();

      if (p.debug_flags & 4) {
        // triage: drain nothing
      } else if (p.mode == kModeDense) {
        float r_i = 0.f, w_i = 1.f;
        if (p.dense_lb && t.valid) {
          r_i = fmaf(rho, st_an, st_c) * 1.000001f;
          if (p.normalize) w_i = 1.f / fmaxf(st_norm, 1e-30f);  // reciprocal: dense_chunk multiplies
        }
        for (int c0 = half * 32; c0 < n32; c0 += 64) dense_chunk<32>(p, t, c0, qc, r_i, w_i);
        if (n32 < p.npad && ((n32 >> 5) & 1) == half) dense_chunk<16>(p, t, n32, qc, r_i, w_i);
      } else if (p.mode == kModeArgmax) {
        // The two warps of a lane quarter share the same 32 rows: warp `half` reduces the 32-column words
        // cb = half, half + 2, ... and they meet through shared memory (named barrier 1 + quarter, 64 threads).
        uint32_t* xq = xchg_base + (warp & 3u) * (32 * 8) + lane * 8;  // 8 words per row
        const uint32_t bar_id = 1u + (warp & 3u);
        uint32_t best = 0, second = 0;
        for (int c0 = half * 32; c0 < n32; c0 += 64) argmax_chunk<32>(p, t, c0, best, second);
        if (n32 < p.npad && ((n32 >> 5) & 1) == half) argmax_chunk<16>(p, t, n32, best, second);
        if (half == 1) { xq[0] = best; xq[1] = second; }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        float thr = 0.f;
        bool flagged = false;
        if (half == 0) {
          const uint32_t b1 = xq[0], s1k = xq[1];
          second = max(max(second, s1k), min(best, b1));
          best = max(best, b1);
          const float r_i = fmaf(rho, st_an, st_c);
          // eps_i = r_i * bn_max bounds |s~ - s| for every query; 2^-15 relative is lost by the key
          const float tol = (2.f * r_i + 6.2e-5f * st_an) * bn_max * 1.0001f;
          const float s1 = ord2f(best & 0xFFFFFF00u);
          const float s2 = ord2f(second & 0xFFFFFF00u);
          thr = s1 - tol;
          flagged = t.valid && second != 0u && (s2 >= thr);
          if (t.valid) p.argmax_out[t.row] = 255 - static_cast<int32_t>(best & 0xFFu);
          xq[2] = __float_as_uint(thr);
          xq[3] = __ballot_sync(0xffffffffu, flagged);
        }
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const uint32_t fl = xq[3];  // same word for the whole quarter: written by every lane of warp `half 0`
        if (fl != 0u) {
          // second pass over the accumulator: bitmask of every query inside the band
          if (half == 1) thr = __uint_as_float(xq[2]);
          uint32_t m[kFlagWords / 2];
#pragma unroll
          for (int i = 0; i < kFlagWords / 2; ++i) {
            m[i] = 0u;
            const int c0 = (2 * i + half) * 32;
            if (c0 + 32 <= p.npad) m[i] = mask_chunk<32>(t, c0, thr);
            else if (c0 < p.npad) m[i] = mask_chunk<16>(t, c0, thr);
            const int nv = p.nq - c0;  // valid columns in this word
            const uint32_t vm = nv >= 32 ? 0xffffffffu : (nv > 0 ? ((1u << nv) - 1u) : 0u);
            m[i] &= vm;
          }
          if (half == 1) { xq[4] = m[0]; xq[5] = m[1]; xq[6] = m[2]; xq[7] = m[3]; }
          asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
          if (half == 0) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(p.flag_count, static_cast<uint32_t>(__popc(fl)));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (flagged) {
              const uint32_t slot = base + __popc(fl & ((1u << lane) - 1u));
              if (slot < p.flag_cap) {
                p.flag_rows[slot] = static_cast<uint32_t>(t.row);
                uint4* dst = reinterpret_cast<uint4*>(p.flag_masks + static_cast<size_t>(slot) * kFlagWords);
                dst[0] = make_uint4(m[0], xq[4], m[1], xq[5]);   // words 0..3 = (half0, half1) interleaved
                dst[1] = make_uint4(m[2], xq[6], m[3], xq[7]);
              }
            }
          }
          // xq is rewritten in the next tile only after both warps pass its first barrier again
        }
      } else {  // kModeThresh
        // invalid rows (past the end of the map): ri = NaN makes every upper bound NaN, and NaN >= T is false for
        // EVERY threshold -- -inf would pass T = -inf, which is what a tiny map (fewer sample groups than k) gets
        float ri = __int_as_float(0x7fc00000), iw = 1.f;
        if (t.valid) {
          ri = fmaf(rho, st_an, st_c);
          if (p.normalize) {
            iw = 1.f / fmaxf(st_norm, 1e-30f);
            ri *= iw;
          }
        }
        if (p.normalize) pend = thresh_tile<true>(p, t, qc, qchunk, iw, ri, half, ring, pend, lane);
        else pend = thresh_tile<false>(p, t, qc, qchunk, iw, ri, half, ring, pend, lane);
      }

      // accumulator stage drained: hand it back to the MMA issuer (leader CTA's barrier); one arrive
      // per warp -- per-thread remote arrives serialise on the barrier (see sim_screen_ts.cu)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) ptx::mbar_arrive_cluster(ptx::smem_u32(bar_tempty + as), 0);
        else ptx::mbar_arrive(ptx::smem_u32(bar_tempty + as));
      }
    }
    if (p.mode == kModeThresh && pend > 0u) pend = ring_flush(p, ring, pend, pend, lane);
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
#endif
}

}  // namespace

static bool stream_b() {  // opt-in variant: B's k-blocks travel with A's instead of staying resident (see the kernel)
  static const bool on = [] { const char* e = getenv("AVL_STREAM_B"); return e && e[0] == '1'; }();
  return on;
}

size_t screen_smem_bytes(int cta_group, int npad, int kblocks, int stages) {
  const size_t bblk = static_cast<size_t>(npad / cta_group) * 128u;
  const size_t b_bytes = bblk * kblocks;
  size_t total = stream_b() ? static_cast<size_t>(stages) * (kStageBytes + bblk)
                            : ((b_bytes + 1023u) & ~size_t(1023)) + static_cast<size_t>(stages) * kStageBytes;
  total += kCtrlBytes + kQConstBytes + kRingBytes + kXchgBytes + 1024u /* alignment slack */;
  // > half of the SM's shared memory, so exactly one CTA (and one 512-column TMEM owner) per SM
  if (total < 120u * 1024u) total = 120u * 1024u;
  return total;
}

int screen_pick_stages(int cta_group, int npad, int kblocks) {
  const size_t limit = 227u * 1024u;
  for (int s = kMaxStages; s >= 2; --s)
    if (screen_smem_bytes(cta_group, npad, kblocks, s) <= limit) return s;
  return 0;
}

int launch_screen(int cta_group, const void* tmap_a, const void* tmap_b, const ScreenParams& p,
                  int num_sms, size_t smem_bytes, cudaStream_t stream) {
  if (p.num_tiles <= 0) return AVL_OK;
  const CUtensorMap& ta = *reinterpret_cast<const CUtensorMap*>(tmap_a);
  const CUtensorMap& tb = *reinterpret_cast<const CUtensorMap*>(tmap_b);
  int units = num_sms / cta_group;
  if (units > p.num_tiles) units = p.num_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(units * cta_group));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cta_group);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (stream_b()) {  // opt-in variant
    auto kernel = cta_group == 2 ? screen_kernel<2, true> : screen_kernel<1, true>;
    AVL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
    AVL_CUDA(cudaLaunchKernelEx(&cfg, kernel, ta, tb, p));
  } else if (cta_group == 2) {
    AVL_CUDA(cudaFuncSetAttribute(screen_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes)));
    AVL_CUDA(cudaLaunchKernelEx(&cfg, screen_kernel<2, false>, ta, tb, p));
  } else {
    AVL_CUDA(cudaFuncSetAttribute(screen_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes)));
    AVL_CUDA(cudaLaunchKernelEx(&cfg, screen_kernel<1, false>, ta, tb, p));
  }
  return AVL_OK;
}

}  // namespace avl
