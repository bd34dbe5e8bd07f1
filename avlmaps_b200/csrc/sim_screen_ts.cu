// Query-stationary variant of the similarity screen for large query batches (128 < Q <= 256,
// D <= 512): the QUERIES are the UMMA A operand and live in TENSOR MEMORY for the whole kernel
// (tcgen05.mma with A from TMEM), the voxel rows are the B operand streamed through shared memory.
//
// Why: in sim_screen.cu (A and B from shared memory) a 256 x 512 query block takes 128 KiB of each
// SM's shared memory, leaving 5 x 16 KiB pipeline stages -- measured on B200: TMA stream alone
// 0.61 ms, MMAs alone 0.57 ms, together 0.80 ms, i.e. the ring is too shallow to overlap them.
// With the queries in TMEM (256 columns: 128 lanes x 512 bf16) all ~200 KiB of shared memory become
// 24 stages of voxel tiles, and the MMAs no longer read A from shared memory at all.
//
//   TMEM columns   [0,256)   A = queries (lane = query, 2 bf16 per column)
//                  [256,384) D stage 0,  [384,512) D stage 1   (128 queries x 128 voxels fp32 each)
//   cta_group::2:  UMMA M = 256 queries (128 per CTA), N = 128 voxels per tile (64 rows of B per CTA)
//   epilogue:      one QUERY per thread, looping over voxel columns; per-voxel error terms come from
//                  a per-warp shared-memory broadcast.
#include <cuda.h>

#include "avl_internal.h"
#include "ptx_sm100.cuh"

namespace avl {
namespace {

constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kMaxStagesTs = 12;
constexpr int kAtomsPerStage = 2;           // k-blocks (64 bf16 = one 128-byte swizzle row) per pipeline stage
constexpr int kBarSlots = 32;
constexpr int kTsTileVox = 128;             // UMMA N: voxels per tile
constexpr int kTsRowsPerCta = kTsTileVox / 2;
constexpr int kTsAtomBytes = kTsRowsPerCta * kBlockK * 2;   // 8 KiB: 64 voxel rows x 64 k
constexpr int kTsStageBytes = kTsAtomBytes * 2;             // 16 KiB per stage (kAtomsPerStage atoms)
constexpr int kCtrlBytes = 1024;
constexpr int kRingEntries = 64;
constexpr int kRingBytes = kEpiWarps * kRingEntries * 12;
constexpr int kStatBytes = kEpiWarps * 2 * 32 * 4;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccBase = 256, kAccStride = 128;

struct Ring {
  uint32_t* row;
  uint32_t* q;
  float* val;
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

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem]^T, 2-CTA form (A: 128 lanes per CTA)
__device__ __forceinline__ void umma_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
screen_ts_kernel(const __grid_constant__ CUtensorMap tmap_v, const ScreenParams p) {
#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
  constexpr int CG = 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;

  uint8_t* smem_v = smem;
  uint8_t* ctrl = smem_v + static_cast<uint32_t>(p.stages) * kTsStageBytes;
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(ctrl);  // [kBarSlots]
  uint64_t* bar_empty = bar_full + kBarSlots;              // [kBarSlots]
  uint64_t* bar_tfull = bar_empty + kBarSlots;             // [2]
  uint64_t* bar_tempty = bar_tfull + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  uint8_t* ring_base = ctrl + kCtrlBytes;
  float* stat_base = reinterpret_cast<float*>(ring_base + kRingBytes);

  ptx::cluster_sync_all();
  if (warp == 0 && lane == 0) ptx::prefetch_tensormap(&tmap_v);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(ptx::smem_u32(bar_full + i), CG);
      ptx::mbar_init(ptx::smem_u32(bar_empty + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(bar_tfull + i), 1);
      ptx::mbar_init(ptx::smem_u32(bar_tempty + i), CG * kEpiWarps);  // one arrive per epilogue warp of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<CG>(ptx::smem_u32(tmem_slot), kTmemCols);
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- queries -> TMEM (A operand): lane = query, column c holds elements (2c, 2c+1)
  if (warp >= 4) {
    const uint32_t lane_base = (warp & 3u) * 32u;
    const uint32_t half = (warp - 4u) >> 2;
    const uint32_t qrow = rank * 128u + lane_base + lane;  // bq has 256 zero-padded rows
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.bq + static_cast<size_t>(qrow) * (p.kblocks * kBlockK));
    const int ncol = p.kblocks * (kBlockK / 2);            // 32-bit columns of A
    for (int c0 = static_cast<int>(half) * 32; c0 < ncol; c0 += 64) {
      uint32_t v[32];
      const uint4* s4 = reinterpret_cast<const uint4*>(src + c0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 t = __ldg(s4 + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
      tmem_st32(tmem_base + (lane_base << 16) + static_cast<uint32_t>(c0), v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();

  const int num_units = static_cast<int>(gridDim.x) / CG;
  const int unit = static_cast<int>(blockIdx.x) / CG;

  if (warp == 0) {
    // ===================== TMA producer: 64 voxel rows x 64 k per stage and CTA =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int j = unit; j < p.num_tiles; j += num_units) {
        const int64_t row0 = static_cast<int64_t>(j) * p.tile_stride * kTsTileVox + rank * kTsRowsPerCta;
        for (int kb = 0; kb < p.kblocks; kb += kAtomsPerStage) {
          const int na = min(kAtomsPerStage, p.kblocks - kb);
          ptx::mbar_wait(ptx::smem_u32(bar_empty + stage), phase ^ 1u, p.dbg, 0x110u + stage);
          if (!(p.debug_flags & 2)) {
            for (int a = 0; a < na; ++a) {
              // tile-major copy (map_prepare_kernel): the 64-row half of tile row0 / 128, k-block kb + a
              const int32_t c0 = p.a_tiled ? 0 : (kb + a) * kBlockK;
              const int32_t c1 = p.a_tiled ? static_cast<int32_t>((((row0 >> 7) * p.kblocks + kb + a) << 7) + (row0 & 127))
                                           : static_cast<int32_t>(row0);
              ptx::tma_load_2d<CG>(ptx::smem_u32(smem_v + stage * kTsStageBytes + a * kTsAtomBytes), &tmap_v,
                                   ptx::smem_u32(bar_full + stage), c0, c1, ptx::kEvictFirst);
            }
            if (leader) ptx::mbar_arrive_expect_tx(ptx::smem_u32(bar_full + stage), kTsAtomBytes * na * CG);
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
    // ===================== MMA issuer =====================
    if (leader && lane == 0) {
      const uint32_t idesc = p.op_f16 ? ptx::make_idesc_f16(256, kTsTileVox) : ptx::make_idesc_bf16(256, kTsTileVox);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int j = unit; j < p.num_tiles; j += num_units, ++it) {
        const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
        ptx::mbar_wait(ptx::smem_u32(bar_tempty + as), aphase ^ 1u, p.dbg, 0x130u + as);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + kAccBase + as * kAccStride;
        for (int kb = 0; kb < p.kblocks; kb += kAtomsPerStage) {
          const int na = min(kAtomsPerStage, p.kblocks - kb);
          ptx::mbar_wait(ptx::smem_u32(bar_full + stage), phase, p.dbg, 0x140u + stage);
          ptx::tc_fence_after();
          for (int a = 0; a < na; ++a) {
            const uint64_t b0 =
                ptx::make_kmajor_sw128_desc(ptx::smem_u32(smem_v + stage * kTsStageBytes + a * kTsAtomBytes));
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)  // 16 bf16 of A = 8 TMEM columns; 32 bytes of the B swizzle row
              if (!(p.debug_flags & 1))
              umma_ts_2cta(tmem_d, tmem_base + static_cast<uint32_t>((kb + a) * (kBlockK / 2) + k * 8), b0 + 2u * k,
                           idesc, ((kb + a) | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit<CG>(ptx::smem_u32(bar_empty + stage));
          if (kb + na >= p.kblocks) ptx::umma_commit<CG>(ptx::smem_u32(bar_tfull + as));
          if (++stage == static_cast<uint32_t>(p.stages)) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: one query per thread, voxels along the columns =====================
    const uint32_t ew = warp - 4u;
    const uint32_t lane_base = (warp & 3u) * 32u;
    const int half = static_cast<int>(ew >> 2);             // columns [64*half, 64*half + 64)
    const int q = static_cast<int>(rank * 128u + lane_base + lane);
    const bool q_valid = q < p.nq;
    const float rho = p.q_glob[0];
    float t_q = __int_as_float(0x7f800000), bn_q = 0.f;     // padded query: threshold +inf
    if (q_valid) {
      bn_q = p.q_bn[q];
      if (p.mode == kModeThresh && !(p.debug_flags & 3)) t_q = p.thr_t[q];  // triage modes: garbage scores, emit nothing
    }
    Ring ring;
    {
      uint8_t* rb = ring_base + ew * (kRingEntries * 12);
      ring.row = reinterpret_cast<uint32_t*>(rb);
      ring.q = ring.row + kRingEntries;
      ring.val = reinterpret_cast<float*>(ring.q + kRingEntries);
    }
    float* st_r = stat_base + ew * 64;   // per-warp broadcast staging: r_j (or r_j/w_j)
    float* st_w = st_r + 32;             //                             1/w_j (thresh) or w_j (dense_lb)
    uint32_t pend = 0, it = 0;

    // per-voxel terms of this lane's voxel in the two 32-voxel chunks of a tile; loaded one tile ahead
    // so that their global-memory latency never sits between the accumulator and its release
    auto load_stats = [&](int jt, float (&sr)[2], float (&sw)[2]) {
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        float r = -INFINITY, w = 1.f;  // voxels past the map: never a candidate / -inf lower bound
        if (jt < p.num_tiles) {
          const int64_t row = static_cast<int64_t>(jt) * p.tile_stride * kTsTileVox + half * 64 + cc * 32 + lane;
          if (row < p.n_rows) {
            r = fmaf(rho, p.row_an[row], p.row_c[row]);
            if (p.normalize) w = fmaxf(p.row_norm[row], 1e-30f);
          }
        }
        const float iw = 1.f / w;
        if (p.mode == kModeThresh) {
          // past the map: NaN, because NaN >= T is false for every threshold (-inf would pass T = -inf)
          sr[cc] = (r == -INFINITY) ? __int_as_float(0x7fc00000) : r * iw;
          sw[cc] = iw;
        } else {  // dense: keep r (slightly inflated for the lower bound) and 1 / w
          sr[cc] = r * 1.000001f;
          sw[cc] = iw;
        }
      }
    };
    float nr[2], nw[2];
    load_stats(unit, nr, nw);

    for (int j = unit; j < p.num_tiles; j += num_units, ++it) {
      const uint32_t as = it & 1u, aphase = (it >> 1) & 1u;
      const int64_t vox0 = static_cast<int64_t>(j) * p.tile_stride * kTsTileVox;  // first voxel of the tile
      const int64_t cvox0 = static_cast<int64_t>(j) * kTsTileVox;                 // compact (sampled launches)
      const float cr[2] = {nr[0], nr[1]}, cw[2] = {nw[0], nw[1]};
      load_stats(j + num_units, nr, nw);  // next tile of this unit, in flight during this tile
      ptx::mbar_wait(ptx::smem_u32(bar_tfull + as), aphase, p.dbg, 0x150u + as);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + kAccBase + as * kAccStride + (lane_base << 16);
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c0 = half * 64 + cc * 32;
        uint32_t v[32];
        ptx::tmem_ld32(taddr + static_cast<uint32_t>(c0), v);
        st_r[lane] = cr[cc];
        st_w[lane] = cw[cc];
        ptx::tmem_ld_wait();
        __syncwarp();
        if (cc == 1) {
          // both chunks of this warp are in registers: hand the accumulator stage back before the math.
          // ONE arrive per warp: 512 per-thread remote arrives per tile serialised on the leader's
          // barrier and cost more than the tile's MMAs (measured: 0.56 ms of pure protocol per pass).
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(ptx::smem_u32(bar_tempty + as), 0);
        }
        if (p.debug_flags & 4) {
          // triage: no epilogue math
        } else if (p.mode == kModeThresh) {
          uint32_t m = 0;
          const float4* r4 = reinterpret_cast<const float4*>(st_r);
          const float4* w4 = reinterpret_cast<const float4*>(st_w);
          if (p.normalize) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 rr = r4[i], ww = w4[i];
              m |= (fmaf(rr.x, bn_q, __uint_as_float(v[4 * i]) * ww.x) >= t_q ? 1u : 0u) << (4 * i);
              m |= (fmaf(rr.y, bn_q, __uint_as_float(v[4 * i + 1]) * ww.y) >= t_q ? 1u : 0u) << (4 * i + 1);
              m |= (fmaf(rr.z, bn_q, __uint_as_float(v[4 * i + 2]) * ww.z) >= t_q ? 1u : 0u) << (4 * i + 2);
              m |= (fmaf(rr.w, bn_q, __uint_as_float(v[4 * i + 3]) * ww.w) >= t_q ? 1u : 0u) << (4 * i + 3);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 rr = r4[i];
              m |= (fmaf(rr.x, bn_q, __uint_as_float(v[4 * i])) >= t_q ? 1u : 0u) << (4 * i);
              m |= (fmaf(rr.y, bn_q, __uint_as_float(v[4 * i + 1])) >= t_q ? 1u : 0u) << (4 * i + 1);
              m |= (fmaf(rr.z, bn_q, __uint_as_float(v[4 * i + 2])) >= t_q ? 1u : 0u) << (4 * i + 2);
              m |= (fmaf(rr.w, bn_q, __uint_as_float(v[4 * i + 3])) >= t_q ? 1u : 0u) << (4 * i + 3);
            }
          }
          // rare: append (voxel row, this thread's query, s~) through the shared-memory ring
          uint32_t u = __reduce_or_sync(0xffffffffu, m);
          while (u) {  // warp-uniform loop over the voxel columns any lane (query) marked
            const int b = __ffs(u) - 1;
            u &= u - 1;
            // the score of column b sits in register v[b] of every lane: select it without dynamic indexing
            float sv = 0.f;
#pragma unroll
            for (int t = 0; t < 32; ++t) sv = (t == b) ? __uint_as_float(v[t]) : sv;
            const bool mine = (m >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, mine);
            if (mine) {
              const uint32_t slot = pend + __popc(bal & ((1u << lane) - 1u));
              ring.row[slot] = static_cast<uint32_t>(vox0 + c0 + b);
              ring.q[slot] = static_cast<uint32_t>(q);
              ring.val[slot] = sv;
            }
            pend += __popc(bal);
            __syncwarp();
            if (pend >= 32u) pend = ring_flush(p, ring, pend, 32u, lane);
          }
        } else if (q_valid && q < p.dense_cols) {  // kModeDense
          float* o = p.dense_out + (cvox0 + c0) * p.dense_rs + static_cast<int64_t>(q) * p.dense_cs;
          if (p.dense_lb) {
#pragma unroll
            for (int t = 0; t < 32; ++t) {
              const float r = st_r[t], iw = st_w[t];  // iw = 1 / w for dense_lb too
              const float x = (__uint_as_float(v[t]) - r * bn_q) * iw;
              o[t * p.dense_rs] = (r == -INFINITY) ? -INFINITY : fmaf(-fabsf(x), 9.5367431640625e-7f, x);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 32; ++t)
              if (vox0 + c0 + t < p.n_rows) o[t * p.dense_rs] = __uint_as_float(v[t]);
          }
        }
        __syncwarp();  // st_r / st_w are rewritten by the next chunk
      }
    }
    if (p.mode == kModeThresh && pend > 0u) pend = ring_flush(p, ring, pend, pend, lane);
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) ptx::tmem_dealloc<CG>(tmem_base, kTmemCols);
#endif
}

}  // namespace

size_t screen_ts_smem_bytes(int stages) {
  return static_cast<size_t>(stages) * kTsStageBytes + kCtrlBytes + kRingBytes + kStatBytes + 1024u;
}
int screen_ts_pick_stages() { return kMaxStagesTs; }

int launch_screen_ts(const void* tmap_v, const ScreenParams& p, int num_sms, size_t smem_bytes, cudaStream_t stream) {
  if (p.num_tiles <= 0) return AVL_OK;
  const CUtensorMap& tv = *reinterpret_cast<const CUtensorMap*>(tmap_v);
  int units = num_sms / 2;
  if (units > p.num_tiles) units = p.num_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(units * 2));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  AVL_CUDA(cudaFuncSetAttribute(screen_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem_bytes)));
  AVL_CUDA(cudaLaunchKernelEx(&cfg, screen_ts_kernel, tv, p));
  return AVL_OK;
}

}  // namespace avl
