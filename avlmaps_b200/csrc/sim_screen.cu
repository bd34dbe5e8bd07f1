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
// Warp roles (384 threads): warps 0-7 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31, i.e. one voxel row per thread;
// the two warps of a lane quarter split the 32-column words), 8 = TMEM allocator, 9 = idle,
// 10 = TMA producer (the leader CTA's is also the tile scheduler), 11 = MMA issuer.  The single-thread roles sit on the HIGHEST warp ids on purpose: the issue
// arbiter of an SM sub-partition prefers the higher warp id, and the producer / MMA issuer are the latency-critical
// instruction streams -- they must not queue behind the ALU-heavy epilogue warps they share a scheduler with.
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
constexpr uint32_t kWarpAlloc = 8, kWarpProducer = 10, kWarpMma = 11;   // warp 9 idles
constexpr int kSchedSlots = 8;     // ring of published tile ids; the producer is never 3 tiles ahead of the epilogue
constexpr int kMaxStages = 12;
constexpr int kCtrlBytes = 512;    // barriers + tmem slot + producer progress word, chunk summaries at +256
constexpr int kQConstBytes = 2048; // float2[256]
constexpr int kXchgBytes = 4 * 32 * 8 * 4;  // argmax only: the two warps of a lane quarter exchange keys / mask words
constexpr int kBucketCntBytes = AVL_MAX_QUERIES * 4;  // threshold mode only: per-query fill counts of this CTA's buckets
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

// Slow path (~2 marked columns per warp and tile).  Kept tiny on purpose: an unrolled per-bit version thrashed the
// instruction cache.  Candidates go to PER-CTA buckets: bucket (query q, CTA c) is `cand_bucket` entries at
// cand_row / cand_val[(q * gridDim.x + c) * cand_bucket], its fill count lives in this CTA's shared memory (one ATOMS
// by lane 0 per marked column, ~60 cycles) and is written to cand_cnt[c][q] once, when the CTA is done.  No global
// atomic is left in the epilogue.  History, all measured on B200: a global atomicAdd per candidate column cost ~2000
// cycles between the accumulator and its release (epilogue 3x slower than the MMAs); staging in a per-warp ring and
// appending 32 at a time hid most of it, but with 16 epilogue warps per CTA pair more than half of the tiles still had
// one warp waiting on a burst, and the accumulator is only released when ALL of them are done; deferring the use of
// the returned slot through registers did not help either (ptxas puts the atomics on one scoreboard, so waiting for the
// oldest waits for the newest).  The finalize kernel gathers a query's buckets (sim_exact.cu).
__device__ __forceinline__ void thresh_emit_word(const ScreenParams& p, uint32_t taddr, int64_t row, int c0,
                                                 uint32_t m, uint32_t* bucket_cnt, uint32_t lane) {
  uint32_t u = __reduce_or_sync(0xffffffffu, m);
  while (u) {  // warp-uniform loop over the columns any lane marked
    const int j = __ffs(u) - 1;
    u &= u - 1;
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr + c0 + j) : "memory");
    ptx::tmem_ld_wait();
    const bool mine = (m >> j) & 1u;
    const uint32_t b = __ballot_sync(0xffffffffu, mine);
    const uint32_t q = static_cast<uint32_t>(c0 + j);
    uint32_t base = 0u;
    if (lane == 0) base = atomicAdd(bucket_cnt + q, static_cast<uint32_t>(__popc(b)));
    base = __shfl_sync(0xffffffffu, base, 0);
    const uint32_t slot = base + __popc(b & ((1u << lane) - 1u));
    if (mine && slot < p.cand_bucket) {
      const size_t o = (static_cast<size_t>(q) * gridDim.x + blockIdx.x) * p.cand_bucket + slot;
      p.cand_row[o] = static_cast<uint32_t>(row);
      p.cand_val[o] = __uint_as_float(v);
    }
  }
}

// One epilogue warp handles the 32-column words cb = half, half + 2, ... of its 32 rows.
template <bool kNorm>
__device__ __forceinline__ void thresh_tile(const ScreenParams& p, const TileCtx& t, const float2* qc,
                                            const float2* qchunk, float iw, float ri, int half,
                                            uint32_t* bucket_cnt, uint32_t lane) {
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
    // deliberately ONE copy of the emission loop (not unrolled): eight inlined copies (4 words x 2 template
    // instances) would crowd the instruction cache the fast path lives in
#pragma unroll 1
    for (int i = 0; i < kFlagWords / 2; ++i) {
      const int c0 = (2 * i + half) * 32;
      const uint32_t mi = i == 0 ? m[0] : (i == 1 ? m[1] : (i == 2 ? m[2] : m[3]));
      if (c0 < p.npad) thresh_emit_word(p, t.taddr, t.row, c0, mi, bucket_cnt, lane);
    }
  }
}

template <int CG>
__global__ void __launch_bounds__(kThreads, 1)
screen_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
              const ScreenParams p) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  // 128B-swizzled tiles need 1024-byte alignment.  The dynamic window starts 1024-aligned (no static shared memory in
  // this kernel); there is no slack to realign with -- at 256 resident queries every KiB is a pipeline stage's -- so a
  // misaligned base is a trap with a watchdog record, not a silent shift.
  extern __shared__ __align__(1024) uint8_t smem[];

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if constexpr (CG == 2) rank = ptx::cluster_ctarank();
  const bool leader = (rank == 0);

  const uint32_t b_rows = static_cast<uint32_t>(p.npad) / CG;
  const uint32_t bblk = b_rows * 128u;                       // bytes of one resident k-block of B
  const uint32_t b_bytes = bblk * static_cast<uint32_t>(p.kblocks);
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + ((b_bytes + 1023u) & ~1023u);
  uint8_t* ctrl = smem_a + static_cast<uint32_t>(p.stages) * kStageBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(ctrl);    // [kMaxStages]
  uint64_t* bar_empty = bar_full + kMaxStages;               // [kMaxStages]
  uint64_t* bar_tfull = bar_empty + kMaxStages;              // [2]
  uint64_t* bar_tempty = bar_tfull + 2;                      // [2]
  uint64_t* bar_bfull = bar_tempty + 2;                      // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_bfull + 1);
  float2* qchunk = reinterpret_cast<float2*>(ctrl + 256);    // [8] per 32-query chunk: (min threshold, max ||b||)
  uint64_t* bar_sched = reinterpret_cast<uint64_t*>(ctrl + 320);  // [kSchedSlots] tile id of iteration it published
  int32_t* sched_tile = reinterpret_cast<int32_t*>(ctrl + 320 + 8 * kSchedSlots);  // [kSchedSlots]
  float2* qc = reinterpret_cast<float2*>(ctrl + kCtrlBytes);
  // argmax mode: exchange words of the two warps of a lane quarter; threshold mode: this CTA's bucket fill counts
  uint32_t* xchg_base = reinterpret_cast<uint32_t*>(ctrl + kCtrlBytes + kQConstBytes);
  uint32_t* bucket_cnt = xchg_base;                          // [AVL_MAX_QUERIES]

  if ((ptx::smem_u32(smem) & 1023u) != 0u) {
    if (threadIdx.x == 0 && p.dbg) {
      p.dbg[0] = 0xDEAD0000u | 0xA11u;
      p.dbg[1] = blockIdx.x;
      p.dbg[2] = ptx::smem_u32(smem);
      __threadfence_system();
    }
    __trap();
  }

  if constexpr (CG == 2) ptx::cluster_sync_all();  // both CTAs resident before the paired TMEM alloc

  if (warp == kWarpProducer && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
  }
  if (warp == kWarpMma && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(bar_full + i), CG);   // leader's expect_tx arrive (+ peer's arrive)
      ptx::mbar_init(ptx::smem_u32(bar_empty + i), 1);   // one tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(bar_tfull + i), 1);          // one tcgen05.commit
      ptx::mbar_init(ptx::smem_u32(bar_tempty + i), CG * kEpiWarps);  // one arrive per epilogue warp of the pair
    }
    ptx::mbar_init(ptx::smem_u32(bar_bfull), CG);
    for (int i = 0; i < kSchedSlots; ++i) ptx::mbar_init(ptx::smem_u32(bar_sched + i), 1);  // one publish per phase
    ptx::fence_barrier_init();
  }
  if (warp == kWarpAlloc) ptx::tmem_alloc<CG>(ptx::smem_u32(tmem_slot), kTmemCols);
  // everything above is independent of the kernels before this one in the stream; from here on their results are read
  // (queries, thresholds) and their scratch is written (candidate buckets, sample maxima)
  pdl_wait();
  pdl_launch_dependents();
  // per-query constants of the threshold screen
  const bool live_thr = p.mode == kModeThresh && !(p.debug_flags & 19);  // triage modes emit nothing
  if (threadIdx.x < 256) {
    const int q = threadIdx.x;
    float2 c = make_float2(__int_as_float(0x7f800000), 0.f);  // +inf: padded column never passes
    if (q < p.nq) {
      c.x = live_thr ? p.thr_t[q] : (p.mode == kModeThresh ? __int_as_float(0x7f800000) : 0.f);
      c.y = p.q_bn[q];
    }
    qc[q] = c;
    if (p.mode == kModeThresh) bucket_cnt[q] = 0u;
  } else if (threadIdx.x < 264) {
    // chunk summaries for the pre-test of the threshold epilogue, straight from global memory (qc is not visible yet)
    const int c0 = (static_cast<int>(threadIdx.x) - 256) * 32;
    float tmin = __int_as_float(0x7f800000), bmax = 0.f;
    if (live_thr) {
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

  // ---- dynamic tile schedule.  Static round-robin left the kernel waiting for its slowest CTA pair: per-pair finishing
  // times spread by 7-15 % on B200 (SMs differ in their distance to the two L2 halves).  The leader CTA's producer
  // takes the next tile from a global counter (one atomic per tile, issued a whole tile ahead of its use) and
  // publishes the id of iteration `it` in slot it % kSchedSlots of BOTH CTAs' shared memory; every role reads it there.
  // The first wave is static (tile = unit); -1 ends the loop.  The counter is never reset: the host advances
  // `tile_base` by num_tiles per launch -- exactly the number of fetches a launch makes.
  auto get_tile = [&](uint32_t it) -> int {
    ptx::mbar_wait_cluster(ptx::smem_u32(bar_sched + (it & (kSchedSlots - 1))), (it / kSchedSlots) & 1u, p.dbg, 0x60u);
    return *reinterpret_cast<volatile int32_t*>(sched_tile + (it & (kSchedSlots - 1)));
  };
  auto publish_tile = [&](uint32_t it, int tile) {  // leader's producer thread only
    const uint32_t slot = it & (kSchedSlots - 1);
    if constexpr (CG == 2)
      ptx::st_and_arrive_cluster(ptx::smem_u32(sched_tile + slot), static_cast<uint32_t>(tile), ptx::smem_u32(bar_sched + slot), 1);
    *reinterpret_cast<volatile int32_t*>(sched_tile + slot) = tile;
    ptx::mbar_arrive(ptx::smem_u32(bar_sched + slot));
  };

  // triage (AVL_DEBUG_FLAGS & 64): SM cycles and nanoseconds of this launch, for the true clock under load
  long long clk0 = 0;
  unsigned long long ns0 = 0;
  if ((p.debug_flags & 64) && threadIdx.x == 0) {
    clk0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns0));
  }

  if (warp == kWarpProducer) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < p.kblocks; ++kb)
        ptx::tma_load_2d<CG>(ptx::smem_u32(smem_b + kb * bblk), &tmap_b, ptx::smem_u32(bar_bfull),
                             kb * kBlockK, static_cast<int32_t>(rank * b_rows), ptx::kEvictLast);
      if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(bar_bfull), b_bytes * CG);
      else ptx::mbar_arrive_cluster(ptx::smem_u32(bar_bfull), 0);

      uint32_t stage = 0, phase = 0, it = 0;
      int j = unit;
      if (leader) publish_tile(0, j);
      for (;; ++it) {
        uint32_t fetched = 0;
        if (leader) fetched = atomicAdd(p.tile_ctr, 1u);   // id of iteration it + 1: not consumed before this tile's loads are out
        else j = get_tile(it);
        if (j < 0) break;
        const int64_t row0 = (static_cast<int64_t>(j) * p.tile_stride * CG + rank) * kTileRows;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          // tile-major copy: the box is rows [(tile * kblocks + kb) * 128, +128) of a 64-element-wide matrix
          const int32_t c0 = p.a_tiled ? 0 : kb * kBlockK;
          const int32_t c1 = p.a_tiled ? static_cast<int32_t>(((row0 >> 7) * p.kblocks + kb) << 7) : static_cast<int32_t>(row0);
          ptx::mbar_wait(ptx::smem_u32(bar_empty + stage), phase ^ 1u, p.dbg, 0x10u + stage);
          if (!(p.debug_flags & 2)) {
            ptx::tma_load_2d<CG>(ptx::smem_u32(smem_a + stage * kStageBytes), &tmap_a,
                                 ptx::smem_u32(bar_full + stage), c0, c1, ptx::kEvictFirst);
            if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(bar_full + stage), kStageBytes * CG);
            else ptx::mbar_arrive_cluster(ptx::smem_u32(bar_full + stage), 0);
          } else {  // triage: barrier protocol only
            if (leader) ptx::mbar_arrive(ptx::smem_u32(bar_full + stage));
            else ptx::mbar_arrive_cluster(ptx::smem_u32(bar_full + stage), 0);
          }
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1u; }
        }
        if (leader) {
          const int nxt = static_cast<int>(fetched - p.tile_base) + num_units;
          j = nxt < p.num_tiles ? nxt : -1;
          publish_tile(it + 1, j);
          if (j < 0) break;
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpMma) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = p.op_f16 ? ptx::make_idesc_f16(kTileRows * CG, static_cast<uint32_t>(p.npad))
                                      : ptx::make_idesc_bf16(kTileRows * CG, static_cast<uint32_t>(p.npad));
      ptx::mbar_wait(ptx::smem_u32(bar_bfull), 0, p.dbg, 0x20u);
      ptx::tc_fence_after();
      uint32_t stage = 0, phase = 0, it = 0;
      for (;; ++it) {
        if (get_tile(it) < 0) break;
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        ptx::mbar_wait(ptx::smem_u32(bar_tempty + as), aphase ^ 1u, p.dbg, 0x30u + as);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * kAccumStride;
        for (int kb = 0; kb < p.kblocks; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(bar_full + stage), phase, p.dbg, 0x40u + stage);
          ptx::tc_fence_after();
          const uint64_t a0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_a + stage * kStageBytes));
          const uint64_t b0 = ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_b + kb * bblk));
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
  } else if (warp < kEpiWarps) {
    // ===================== epilogue (TMEM -> registers -> fused reduction) =====================
    const uint32_t quarter = warp & 3u;                   // TMEM lane quarter this warp may read (hardware: warp id % 4)
    const uint32_t lane_base = quarter * 32u;
    const int half = static_cast<int>(warp >> 2);         // which interleaved set of 32-column words
    const float rho = p.q_glob[0], bn_max = p.q_glob[1];
    uint32_t it = 0;
    for (;; ++it) {
      const int j = get_tile(it);
      if (j < 0) break;
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
      } else if (p.debug_flags & 32) {
        // triage: the TMEM reads of the threshold epilogue without its arithmetic
        for (int c0 = half * 32; c0 < n32; c0 += 64) {
          uint32_t v[32];
          ptx::tmem_ld32(t.taddr + c0, v);
          ptx::tmem_ld_wait();
          asm volatile("" ::"r"(v[0]), "r"(v[31]));
        }
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
        uint32_t* xq = xchg_base + quarter * (32 * 8) + lane * 8;  // 8 words per row
        const uint32_t bar_id = 1u + quarter;
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
        if (p.normalize) thresh_tile<true>(p, t, qc, qchunk, iw, ri, half, bucket_cnt, lane);
        else thresh_tile<false>(p, t, qc, qchunk, iw, ri, half, bucket_cnt, lane);
      }

      // accumulator stage drained: hand it back to the MMA issuer (leader CTA's barrier); one arrive
      // per warp -- per-thread remote arrives serialise on the barrier
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) ptx::mbar_arrive_cluster(ptx::smem_u32(bar_tempty + as), 0);
        else ptx::mbar_arrive(ptx::smem_u32(bar_tempty + as));
      }
    }
    if ((p.debug_flags & 64) && threadIdx.x == 0 && p.dbg && blockIdx.x < 160) p.dbg[16 + 4 * 160 + blockIdx.x] = it;
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all(); else __syncthreads();
  // bucket fill counts of this CTA (may exceed cand_bucket: the finalize kernel flags that query as overflowed)
  if (p.mode == kModeThresh && threadIdx.x < static_cast<uint32_t>(p.nq))
    p.cand_cnt[static_cast<size_t>(blockIdx.x) * AVL_MAX_QUERIES + threadIdx.x] = bucket_cnt[threadIdx.x];
  if (warp == kWarpAlloc) ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
  if ((p.debug_flags & 64) && threadIdx.x == 0 && p.dbg) {
    unsigned long long ns1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns1));
    const long long cyc = clock64() - clk0;
    if (blockIdx.x == 0) {
      p.dbg[8] = static_cast<uint32_t>(cyc);
      p.dbg[9] = static_cast<uint32_t>(cyc >> 32);
      p.dbg[10] = static_cast<uint32_t>(ns1 - ns0);
    }
    if (blockIdx.x < 160) {  // per-CTA record: start (ns, low 32 bits), duration (ns), cycles, SM id
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      uint32_t* r = p.dbg + 16 + blockIdx.x * 4;
      r[0] = static_cast<uint32_t>(ns0);
      r[1] = static_cast<uint32_t>(ns1 - ns0);
      r[2] = static_cast<uint32_t>(cyc);
      r[3] = smid;
    }
  }
#endif
}

}  // namespace

// Shared memory of one CTA: resident B, the A ring, barriers + per-query constants, and -- argmax mode only -- the
// exchange words of the two warps of a lane quarter.  No alignment slack (the kernel checks its base).
size_t screen_smem_bytes(int cta_group, int npad, int kblocks, int stages, int mode) {
  const size_t bblk = static_cast<size_t>(npad / cta_group) * 128u;
  const size_t b_bytes = bblk * kblocks;
  size_t total = ((b_bytes + 1023u) & ~size_t(1023)) + static_cast<size_t>(stages) * kStageBytes;
  total += kCtrlBytes + kQConstBytes + (mode == kModeArgmax ? kXchgBytes : (mode == kModeThresh ? kBucketCntBytes : 0));
  // > half of the SM's shared memory, so exactly one CTA (and one 512-column TMEM owner) per SM
  if (total < 120u * 1024u) total = 120u * 1024u;
  return total;
}

int screen_pick_stages(int cta_group, int npad, int kblocks, int mode) {
  const size_t limit = 227u * 1024u;
  static const int cap = [] {  // A/B: AVL_MAX_STAGES caps the ring depth
    const char* e = getenv("AVL_MAX_STAGES");
    const int v = e ? atoi(e) : kMaxStages;
    return v >= 2 && v <= kMaxStages ? v : kMaxStages;
  }();
  for (int s = cap; s >= 2; --s)
    if (screen_smem_bytes(cta_group, npad, kblocks, s, mode) <= limit) return s;
  return 0;
}

int screen_grid(int cta_group, int num_sms, int num_tiles) {
  int units = num_sms / cta_group;
  if (units > num_tiles) units = num_tiles;
  return units * cta_group;
}

int launch_screen(int cta_group, const void* tmap_a, const void* tmap_b, const ScreenParams& p,
                  int num_sms, size_t smem_bytes, cudaStream_t stream) {
  if (p.num_tiles <= 0) return AVL_OK;
  const CUtensorMap& ta = *reinterpret_cast<const CUtensorMap*>(tmap_a);
  const CUtensorMap& tb = *reinterpret_cast<const CUtensorMap*>(tmap_b);
  const int units = screen_grid(cta_group, num_sms, p.num_tiles) / cta_group;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(units * cta_group));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned>(cta_group);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (cta_group == 2) {
    AVL_CUDA(cudaFuncSetAttribute(screen_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes)));
    AVL_CUDA(cudaLaunchKernelEx(&cfg, screen_kernel<2>, ta, tb, p));
  } else {
    AVL_CUDA(cudaFuncSetAttribute(screen_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(smem_bytes)));
    AVL_CUDA(cudaLaunchKernelEx(&cfg, screen_kernel<1>, ta, tb, p));
  }
  return AVL_OK;
}

}  // namespace avl
