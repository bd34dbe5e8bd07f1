// Exact (fp64-accumulated) kernels of the landmark-index path and the small selection
// kernels around the tcgen05 screen.
//
// Canonical score (what oracle/ computes and what every index entry point returns):
//     dot   = sum_k (double)a[k] * (double)b[k]        (each product exact in fp64)
//     s     = (float)dot
//     s     = s * inv_norm_i        if normalize_map   (inv_norm_i = 1.0f / (float)sqrt(sum a^2), 0 if the row is 0)
//     s     = s * scale_q           if scale given
// which restates `map_feats @ text_feats.T` (reference avlmaps/utils/clip_utils.py:229,240) and
// `scale * audio_features @ text_features.T` (avlmaps/map/sound_map.py:109) with one rounding
// instead of OpenBLAS' unknowable fp32 summation order.
#include <algorithm>
#include <cfloat>

#include <cuda_fp16.h>

#include "avl_internal.h"

namespace avl {

namespace {

__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return b ^ (static_cast<uint32_t>(static_cast<int32_t>(b) >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  uint32_t b = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
  return __uint_as_float(b);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float canon_score(double dot, float inv_norm, int normalize, const float* scale,
                                             int q) {
  float s = static_cast<float>(dot);
  if (normalize) s = __fmul_rn(s, inv_norm);
  if (scale) s = __fmul_rn(s, scale[q]);
  return s;
}
__device__ __forceinline__ float inv_of_norm(float nrm) { return nrm > 0.f ? __fdiv_rn(1.0f, nrm) : 0.f; }

// tensor-core operand element: bf16 (8-bit mantissa, fp32 range) or fp16 (11-bit mantissa: 8x smaller rounding
// residual, hence an 8x tighter error band; range 65504).  Returns the 16 bits and the value they stand for.
__device__ __forceinline__ uint16_t to_operand(float x, int f16, float* back) {
  if (f16) {
    const __half h = __float2half_rn(x);
    *back = __half2float(h);
    return __half_as_ushort(h);
  }
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  *back = __bfloat162float(h);
  return __bfloat16_as_ushort(h);
}

// ------------------------------------------------------------------ map_prepare
// One warp per row: bf16 copy (zero padded to dpad), fp32 norm, rounding residual norms.
// tiled: the operand copy is stored tile-major -- [tile of 128 rows][k-block of 64][row][64 elements] -- so that the
// 16 KiB box one TMA load of the screen kernel fetches (128 rows x 128 bytes) is ONE contiguous 16 KiB run of HBM
// and a whole 128-row tile is one contiguous 128 * dpad * 2 bytes (what its L2 prefetch names with one instruction),
// instead of 128 separate 128-byte pieces 2 * dpad bytes apart.
__global__ void map_prepare_kernel(const float* __restrict__ feat, int64_t n, int32_t d, int32_t dpad,
                                   __nv_bfloat16* __restrict__ bf, float* __restrict__ row_norm,
                                   float* __restrict__ row_c, float* __restrict__ row_an, float kappa, int f16,
                                   uint32_t* __restrict__ nonfinite, int tiled) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* a = feat + row * d;
  uint16_t* o = reinterpret_cast<uint16_t*>(bf + row * dpad);
  const int64_t tile_base = (row / kTileRows) * kTileRows * dpad + (row % kTileRows) * kBlockK;
  double sa = 0.0, sd = 0.0, sb = 0.0;
  bool bad = false;
  for (int k = lane; k < dpad; k += 32) {
    const float x = k < d ? a[k] : 0.f;
    float xb;
    const uint16_t ob = to_operand(x, f16, &xb);
    if (tiled)
      reinterpret_cast<uint16_t*>(bf)[tile_base + static_cast<int64_t>(k / kBlockK) * (kTileRows * kBlockK) + (k % kBlockK)] = ob;
    else
      o[k] = ob;
    bad |= !isfinite(xb);    // |x| beyond the fp16 range (or a non-finite input): the caller falls back to bf16
    const float e = x - xb;  // exact
    sa += static_cast<double>(x) * x;
    sd += static_cast<double>(e) * e;
    sb += static_cast<double>(xb) * xb;
  }
  sa = warp_sum(sa);
  sd = warp_sum(sd);
  sb = warp_sum(sb);
  // fp16 operands: a row whose values sink into fp16's subnormals (a far point of a fused map is alpha * f with alpha
  // down to e^-30) loses more to rounding than bf16 would -- its band then covers every query.  Like a value beyond
  // the fp16 range, such a row makes the map fall back to bf16 operands (same exponent range as fp32).
  if (f16 && sd > sa * 1.6e-5) bad = true;  // relative residual above 2^-8, bf16's worst case
  if (__any_sync(0xffffffffu, bad) && lane == 0 && nonfinite) atomicExch(nonfinite, 1u);
  if (lane == 0) {
    const float an = __double2float_ru(sqrt(sb) * (1.0 + 1e-7));
    row_norm[row] = static_cast<float>(sqrt(sa));
    row_an[row] = an;
    row_c[row] = __double2float_ru(sqrt(sd) * (1.0 + 1e-7) + static_cast<double>(kappa) * an);
  }
}

// ------------------------------------------------------------------ query_prepare
// One block per padded query row: bf16 copy, ||b||, ||b - bf16(b)|| / ||b||.
// glob[0] = max ratio (rho), glob[1] = max ||b||: reduced by the LAST block to finish (ticket; it also re-arms the
// ticket), so no zero-fill precedes the launch -- one stream operation less per query batch.
__global__ void query_prepare_kernel(const float* __restrict__ q, const float* __restrict__ fold_scale, int32_t nq,
                                     int32_t d, int32_t dpad, __nv_bfloat16* __restrict__ bq,
                                     float* __restrict__ q_bn, float* __restrict__ q_ratio, uint32_t* __restrict__ glob,
                                     uint32_t* __restrict__ ticket, int f16) {
  pdl_wait();               // the previous call's finalize / fallback still read q_bn and glob
  pdl_launch_dependents();
  const int r = blockIdx.x;
  __shared__ double red[2][32];
  __shared__ uint32_t sh_last;
  double sb = 0.0, sd = 0.0;
  for (int k = threadIdx.x; k < dpad; k += blockDim.x) {
    float x = (r < nq && k < d) ? q[static_cast<size_t>(r) * d + k] : 0.f;
    // argmax mode compares ACROSS queries, so the per-query scale is folded into B (the bounds below
    // are then those of the scaled vector); top-k mode keeps B unscaled and divides its threshold
    if (fold_scale && r < nq) x = __fmul_rn(x, fold_scale[r]);
    float xb;
    reinterpret_cast<uint16_t*>(bq)[static_cast<size_t>(r) * dpad + k] = to_operand(x, f16, &xb);
    const float e = x - xb;
    sb += static_cast<double>(x) * x;
    sd += static_cast<double>(e) * e;
  }
  sb = warp_sum(sb);
  sd = warp_sum(sd);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) { red[0][w] = sb; red[1][w] = sd; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tb = 0.0, td = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) { tb += red[0][i]; td += red[1][i]; }
    float bn = 0.f, ratio = 0.f;
    if (r < nq) {
      bn = __double2float_ru(sqrt(tb) * (1.0 + 1e-7));
      ratio = (tb > 0.0 ? __double2float_ru(sqrt(td / tb) * (1.0 + 1e-6)) : 0.f) + 2e-6f;
      q_bn[r] = bn;
    }
    q_ratio[r] = ratio;  // padded rows: 0
    __threadfence();
    sh_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (sh_last) {
    __threadfence();
    float mr = 0.f, mb = 0.f;  // both non-negative
    for (int i = threadIdx.x; i < static_cast<int>(gridDim.x); i += blockDim.x) {
      mr = fmaxf(mr, __ldcg(q_ratio + i));
      if (i < nq) mb = fmaxf(mb, __ldcg(q_bn + i));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    }
    __shared__ float smr[32], smb[32];
    if (l == 0) { smr[w] = mr; smb[w] = mb; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int i = 1; i < (blockDim.x >> 5); ++i) { mr = fmaxf(mr, smr[i]); mb = fmaxf(mb, smb[i]); }
      glob[0] = __float_as_uint(mr);
      glob[1] = __float_as_uint(mb);
      *ticket = 0u;
    }
  }
}

// ------------------------------------------------------------------ dense exact scores
// 64 rows x 64 queries per block, 4x4 outputs per thread, fp64 accumulators, k ascending.
constexpr int kDT = 64, kDK = 16;
__global__ void __launch_bounds__(256)
dense_exact_kernel(const float* __restrict__ feat, int64_t n, int32_t d, const float* __restrict__ q,
                   int32_t nq, const float* __restrict__ scale, const float* __restrict__ row_norm,
                   int normalize, float* __restrict__ out, int64_t out_rs, int64_t out_cs) {
  __shared__ float As[kDK][kDT + 4];
  __shared__ float Bs[kDK][kDT + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kDT;
  const int q0 = blockIdx.y * kDT;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < d; k0 += kDK) {
    // 64 x 16 tile of A and of B, one element (x4) per thread, coalesced along k
    for (int e = threadIdx.x; e < kDT * kDK; e += 256) {
      const int r = e / kDK, k = e % kDK;
      const int64_t gr = row0 + r;
      As[k][r] = (gr < n && k0 + k < d) ? feat[gr * d + k0 + k] : 0.f;
      const int gq = q0 + r;
      Bs[k][r] = (gq < nq && k0 + k < d) ? q[static_cast<size_t>(gq) * d + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kDK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t gr = row0 + ty * 4 + i;
    if (gr >= n) continue;
    const float inv = normalize ? inv_of_norm(row_norm[gr]) : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gq = q0 + tx * 4 + j;
      if (gq < nq) out[gr * out_rs + gq * out_cs] = canon_score(acc[i][j], inv, normalize, scale, gq);
    }
  }
}

// warp-cooperative exact dot of fp32 row a (global) with fp32 row b (global).  Lane l sums
// k = l, l+32, ... in ascending order; 512-element chunks are loaded with 32 independent
// requests in flight before the fp64 FMA chain starts (the loop is latency-bound otherwise).
__device__ __forceinline__ double warp_dot(const float* __restrict__ a, const float* __restrict__ b, int d,
                                           int lane) {
  double s = 0.0;
  int k0 = 0;
  for (; k0 + 512 <= d; k0 += 512) {
    float av[16], bv[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      av[t] = __ldg(a + k0 + t * 32 + lane);
      bv[t] = __ldg(b + k0 + t * 32 + lane);
    }
#pragma unroll
    for (int t = 0; t < 16; ++t) s = fma(static_cast<double>(av[t]), static_cast<double>(bv[t]), s);
  }
  for (int k = k0 + lane; k < d; k += 32) s = fma(static_cast<double>(a[k]), static_cast<double>(b[k]), s);
  return warp_sum(s);
}

// ------------------------------------------------------------------ argmax re-rank
// One warp per flagged row: exact scores of the queries in its band mask, first max wins.
// float->double conversions run at a quarter of the FMA rate, so the queries are converted once per
// call (q64) and the row once per flagged row (registers) instead of once per product.
__global__ void f32_to_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = static_cast<double>(src[i]);
}

__global__ void __launch_bounds__(256)
argmax_rerank_kernel(const float* __restrict__ feat, int32_t d, const float* __restrict__ q,
                     const double* __restrict__ q64, const float* __restrict__ q_bn, int32_t nq,
                     const float* __restrict__ scale, const float* __restrict__ row_norm, int normalize,
                     const uint32_t* __restrict__ flag_count,
                     const uint32_t* __restrict__ flag_rows, const uint32_t* __restrict__ flag_masks,
                     uint32_t flag_cap, int32_t* __restrict__ argmax_out) {
  const int lane = threadIdx.x & 31;
  const uint32_t nflag = min(*flag_count, flag_cap);
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t e0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (d == 512) {
    // The fp32 rows of the flagged voxels are scattered 2 KiB reads from HBM: the rows of this warp's next
    // two entries are prefetched into L2 (16 lanes x 128 B each, no registers held) while the candidates
    // of the current one are scored.
    auto prefetch_row = [&](uint32_t e) {
      if (e < nflag && lane < 16) {
        const float* r = feat + static_cast<int64_t>(flag_rows[e]) * 512 + lane * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
      }
    };
    prefetch_row(e0);
    prefetch_row(e0 + nwarps);
    for (uint32_t e = e0; e < nflag; e += nwarps) {
      prefetch_row(e + 2 * nwarps);
      const int64_t row = flag_rows[e];
      float cur[16];
#pragma unroll
      for (int t = 0; t < 16; ++t) cur[t] = __ldg(feat + row * 512 + t * 32 + lane);
      const float inv = normalize ? inv_of_norm(row_norm[row]) : 1.f;
      const uint4 m0 = *reinterpret_cast<const uint4*>(flag_masks + static_cast<size_t>(e) * kFlagWords);
      const uint4 m1 = *reinterpret_cast<const uint4*>(flag_masks + static_cast<size_t>(e) * kFlagWords + 4);
      const uint32_t masks[kFlagWords] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
      // Stage 1: fp32 FMA dots with a rigorous error bound.  |fp32 dot - exact| <= (16 + 5) u sum|a_k b_k| <=
      // 1.3e-6 ||a|| ||b||; the canonical score is within 4u of the exact one.  If the best lower bound beats
      // every other upper bound the argmax is decided without touching fp64 (the usual case).
      const float anorm = row_norm[row] * 1.000001f;
      float bestL = -FLT_MAX, u1 = -FLT_MAX, u2 = -FLT_MAX;
      int qL = -1, qU = -1;
#pragma unroll
      for (int w = 0; w < kFlagWords; ++w) {
        uint32_t m = masks[w];
        while (m) {
          const int bit = __ffs(m) - 1;
          m &= m - 1;
          const int qq = w * 32 + bit;
          const float* b = q + static_cast<size_t>(qq) * 512;
          float bf[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) bf[t] = __ldg(b + t * 32 + lane);
          float sum = 0.f;
#pragma unroll
          for (int t = 0; t < 16; ++t) sum = fmaf(cur[t], bf[t], sum);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          float f = sum, g = 2e-6f * anorm * q_bn[qq];  // q_bn: norm of the scale-folded query (query_prepare)
          if (normalize) { f *= inv; g *= inv; }
          if (scale) f *= scale[qq];
          const float lo = f - g, hi = f + g;
          if (lo > bestL) { bestL = lo; qL = qq; }
          if (hi > u1) { u2 = u1; u1 = hi; qU = qq; } else if (hi > u2) { u2 = hi; }
        }
      }
      int best_q = -1;
      if (qL >= 0 && qL == qU && u2 < bestL) {
        best_q = qL;
      } else {
        // Stage 2 (rare): the fp32 bounds overlap -> canonical fp64-accumulated scores, first maximum wins
        double ad[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) ad[t] = static_cast<double>(cur[t]);
        float best = -FLT_MAX;
#pragma unroll
        for (int w = 0; w < kFlagWords; ++w) {
          uint32_t m = masks[w];
          while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int qq = w * 32 + bit;
            const double* b = q64 + static_cast<size_t>(qq) * 512;
            double bd[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) bd[t] = __ldg(b + t * 32 + lane);
            double sum = 0.0;
#pragma unroll
            for (int t = 0; t < 16; ++t) sum = fma(ad[t], bd[t], sum);  // k ascending per lane, like warp_dot
            const float sc = canon_score(warp_sum(sum), inv, normalize, scale, qq);
            if (best_q < 0 || sc > best) { best = sc; best_q = qq; }
          }
        }
      }
      if (lane == 0 && best_q >= 0) argmax_out[row] = best_q;
    }
    return;
  }
  for (uint32_t e = e0; e < nflag; e += nwarps) {
    const int64_t row = flag_rows[e];
    const float* a = feat + row * d;
    const float inv = normalize ? inv_of_norm(row_norm[row]) : 1.f;
    float best = -FLT_MAX;
    int best_q = -1;
    for (int w = 0; w < kFlagWords; ++w) {
      uint32_t m = flag_masks[static_cast<size_t>(e) * kFlagWords + w];
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        const int qq = w * 32 + bit;
        const double dot = warp_dot(a, q + static_cast<size_t>(qq) * d, d, lane);
        const float sc = canon_score(dot, inv, normalize, scale, qq);
        if (best_q < 0 || sc > best) { best = sc; best_q = qq; }
      }
    }
    if (lane == 0 && best_q >= 0) argmax_out[row] = best_q;
  }
}

// ------------------------------------------------------------------ one exact score column
// out[i] = canonical score of row i against ONE query (warp per row).  Used by the top-k fallback,
// where the 64x64-tile dense kernel would waste 63/64 of its work.
__global__ void __launch_bounds__(256)
column_exact_kernel(const float* __restrict__ feat, int64_t n, int32_t d, const float* __restrict__ q,
                    const float* __restrict__ scale, const float* __restrict__ row_norm, int normalize,
                    float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; row < n; row += nwarps) {
    const double dot = warp_dot(feat + row * d, q, d, lane);
    if (lane == 0) {
      const float inv = normalize ? inv_of_norm(row_norm[row]) : 1.f;
      out[row] = canon_score(dot, inv, normalize, scale, 0);
    }
  }
}

// ------------------------------------------------------------------ block-wide selection helpers
// k-th largest (1-based) of keys[0..n) by bitwise bisection; every thread returns the value.
template <typename KeyT, typename Fetch>
__device__ KeyT block_kth_largest(Fetch fetch, int n, int k, int* sh_cnt) {
  KeyT v = 0;
  constexpr int kBits = sizeof(KeyT) * 8;
  for (int bit = kBits - 1; bit >= 0; --bit) {
    const KeyT cand = v | (static_cast<KeyT>(1) << bit);
    int c = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) c += (fetch(j) >= cand) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x == 0) *sh_cnt = 0;
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(sh_cnt, c);
    __syncthreads();
    if (*sh_cnt >= k) v = cand;
    __syncthreads();
  }
  return v;
}

// Same result for 32-bit keys with 4 passes of an 8-bit radix histogram in shared memory instead of 32 rounds of
// bisection (each round is a block-wide count with three barriers; the finalize kernels spend most of their
// instructions there).  hist: 256 shared words; every thread returns the value.
template <typename Fetch>
__device__ uint32_t block_kth_largest_radix32(Fetch fetch, int n, int k, uint32_t* hist, int* sh_sel) {
  uint32_t prefix = 0u, mask = 0u;
  int kk = k;  // rank still to find among the keys that match the prefix
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0u;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const uint32_t key = fetch(j);
      if ((key & mask) == prefix) atomicAdd(hist + ((key >> shift) & 255u), 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns digits [8l, 8l + 8); walk from the largest digit down until kk keys are covered
      const int lane = threadIdx.x;
      uint32_t c[8], tot = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; tot += c[i]; }
      uint32_t suffix = tot;  // inclusive suffix sums over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_down_sync(0xffffffffu, suffix, o);
        if (lane + o < 32) suffix += y;
      }
      const uint32_t above = suffix - tot;  // keys with a digit in a higher lane
      if (above < static_cast<uint32_t>(kk) && static_cast<uint32_t>(kk) <= above + tot) {
        uint32_t run = above;
        for (int i = 7; i >= 0; --i) {
          if (run + c[i] >= static_cast<uint32_t>(kk)) {
            sh_sel[0] = lane * 8 + i;             // the digit of the k-th largest key
            sh_sel[1] = kk - static_cast<int>(run);  // its rank among the keys with that digit
            break;
          }
          run += c[i];
        }
      }
    }
    __syncthreads();
    prefix |= static_cast<uint32_t>(sh_sel[0]) << shift;
    mask |= 255u << shift;
    kk = sh_sel[1];  // sh_sel is rewritten only after the next pass's two barriers
  }
  return prefix;
}

// ------------------------------------------------------------------ threshold from a sample
// sample_lb[q][c]: LOWER bounds (s~ - eps)/w of the sampled rows, written transposed by the screen
// kernel (dense_lb; rows past the end of the map are -inf).  Per query: T_q = k-th largest of the
// per-thread maxima.  The maxima belong to distinct rows, so T_q <= k-th largest lower bound over the
// sample <= k-th largest exact score over the whole map (in units of score / scale_q): a valid
// threshold from ONE coalesced pass over the sample.
constexpr int kSelThreads = 1024;
__global__ void __launch_bounds__(kSelThreads)
select_threshold_kernel(const float* __restrict__ sample_lb, int32_t n_sample, int64_t ld, int32_t k,
                        float* __restrict__ thr_t) {
  pdl_wait();
  pdl_launch_dependents();
  const int q = blockIdx.x;
  const float* col = sample_lb + static_cast<int64_t>(q) * ld;
  uint32_t best = 0u;  // 0 = this thread saw no valid row
  int c = threadIdx.x;
  for (; c + 3 * kSelThreads < n_sample; c += 4 * kSelThreads) {  // 4 independent loads in flight
    const float a0 = col[c], a1 = col[c + kSelThreads], a2 = col[c + 2 * kSelThreads], a3 = col[c + 3 * kSelThreads];
    const float m = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
    if (m > -INFINITY) best = max(best, max(f2ord(m), 1u));
  }
  for (; c < n_sample; c += kSelThreads) {
    const float a0 = col[c];
    if (a0 > -INFINITY) best = max(best, max(f2ord(a0), 1u));
  }
  const int groups = __syncthreads_count(best != 0u);
  if (groups < k) {
    if (threadIdx.x == 0) thr_t[q] = -INFINITY;  // tiny maps: every row is a candidate
    return;
  }
  uint32_t v = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = v | (1u << bit);
    if (__syncthreads_count(best >= cand) >= k) v = cand;
  }
  if (threadIdx.x == 0) thr_t[q] = ord2f(v);
}

// ------------------------------------------------------------------ top-k finalize
// One block per query.  Gathers its entries from the unified candidate list, finds the k-th best
// LOWER bound, keeps the entries whose UPPER bound reaches it, re-scores those exactly (fp64) and
// orders them by (score desc, row asc).
constexpr int kFinThreads = 512;
constexpr int kFinMaxGrid = 256;   // CTAs of one screen launch (one per SM)
__global__ void __launch_bounds__(kFinThreads, 2)
topk_finalize_kernel(const float* __restrict__ feat, int64_t n_rows, int32_t d, const float* __restrict__ q,
                     const float* __restrict__ scale, const float* __restrict__ row_norm,
                     const float* __restrict__ row_c, const float* __restrict__ row_an,
                     const float* __restrict__ q_bn, const uint32_t* __restrict__ glob, int normalize,
                     int32_t k, const uint32_t* __restrict__ bucket_cnt, int32_t grid, uint32_t cand_bucket,
                     const uint32_t* __restrict__ cand_row, const float* __restrict__ cand_val, uint32_t cand_cap,
                     int64_t* __restrict__ out_idx, float* __restrict__ out_score,
                     uint32_t* __restrict__ cand_total, uint32_t* __restrict__ overflow_flags, uint32_t* gscratch) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ uint8_t sm[];
  // `gscratch` (pipelined calls): the three per-candidate arrays live in global memory (L2-resident, 12 bytes x cand_cap
  // per query) instead of shared memory, and the block has 256 threads: it then fits on an SM NEXT TO a CTA of the
  // persistent screen kernel (which leaves ~15 KiB of shared memory and 22 k registers), so the finalize of one query
  // batch runs while the next batch's screen streams the map -- slower per block, but off the critical path.
  uint32_t* Lk = gscratch ? gscratch + static_cast<size_t>(blockIdx.x) * 3 * cand_cap : reinterpret_cast<uint32_t*>(sm);
  uint32_t* Ix = Lk + cand_cap;
  uint32_t* Uk = Ix + cand_cap;
  // 12 bytes per candidate = 96 KiB at 8192 candidates, 512 threads: TWO blocks per SM, so the 256 queries of a full
  // batch are one wave over 148 SMs instead of two.  The exact keys of the survivors live where the upper bounds were
  // (dead once the survivors are chosen): at most cand_cap / 2 survivors, more flags the query for the exact fallback.
  unsigned long long* K64 = reinterpret_cast<unsigned long long*>(Uk);
  __shared__ int sh_cnt;
  __shared__ int sh_ns;
  __shared__ uint32_t sh_off[kFinMaxGrid + 1];
  __shared__ uint32_t sh_bad;
  const int qq = blockIdx.x;
  // The screen left this query's candidates in one bucket per CTA (sim_screen.cu): exclusive scan of the fill counts,
  // then every warp copies whole buckets.  A bucket that overflowed, or more candidates than this block can hold,
  // flags the query for the exact fallback (adversarial data, massive ties).
  if (threadIdx.x == 0) { sh_ns = 0; sh_bad = 0u; }
  __syncthreads();
  // all fill counts in one round of loads (one per thread), then one warp scans them from shared memory
  for (int c = threadIdx.x; c < grid; c += blockDim.x) {
    const uint32_t v = bucket_cnt[static_cast<size_t>(c) * AVL_MAX_QUERIES + qq];
    if (v > cand_bucket) sh_bad = 1u;
    sh_off[c + 1] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t run = 0u;
    if (threadIdx.x == 0) sh_off[0] = 0u;
    for (int c0 = 0; c0 < grid; c0 += 32) {
      const int c = c0 + static_cast<int>(threadIdx.x);
      const uint32_t v = c < grid ? sh_off[c + 1] : 0u;
      uint32_t x = v;  // inclusive warp scan
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (static_cast<int>(threadIdx.x) >= o) x += y;
      }
      __syncwarp();
      if (c < grid) sh_off[c + 1] = run + x;   // exclusive offset of bucket c + 1 = inclusive sum up to c
      run += __shfl_sync(0xffffffffu, x, 31);
    }
  }
  __syncthreads();
  const uint32_t cnt = sh_off[grid];
  if (threadIdx.x == 0) cand_total[qq] = cnt;
  if (sh_bad || cnt > cand_cap) {
    if (threadIdx.x == 0) overflow_flags[qq] = 1u;
    return;
  }
  if (threadIdx.x == 0) overflow_flags[qq] = 0u;
  // Gather + bounds in ONE pass over the candidates, flattened over the block (a thread finds its candidate's bucket
  // by bisection of the offsets): the loads of the candidate, of its row statistics and the bound arithmetic pipeline
  // instead of forming three latency-bound phases, and no warp walks ten buckets one after the other.
  const int n = static_cast<int>(cnt);
  const float rho = __uint_as_float(glob[0]);
  const float bn = q_bn[qq];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    int lo = 0, hi = grid;          // largest c with sh_off[c] <= j
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (sh_off[mid] <= static_cast<uint32_t>(j)) lo = mid; else hi = mid;
    }
    const size_t src = (static_cast<size_t>(qq) * grid + lo) * cand_bucket + (static_cast<uint32_t>(j) - sh_off[lo]);
    const uint32_t i = cand_row[src];
    const float s = cand_val[src];
    const float r_i = fmaf(rho, row_an[i], row_c[i]);
    const float w_i = normalize ? fmaxf(row_norm[i], 1e-30f) : 1.f;
    const float e = __fmul_ru(r_i, bn);
    Ix[j] = i;
    Lk[j] = max(f2ord(__fdiv_rd(__fsub_rd(s, e), w_i)), 1u);
    Uk[j] = max(f2ord(__fdiv_ru(__fadd_ru(s, e), w_i)), 1u);
  }
  __syncthreads();
  const int kk = min(k, n);
  uint32_t v = 0u;
  __shared__ uint32_t hist[256];
  __shared__ int sh_sel[2];
  if (kk > 0) v = block_kth_largest_radix32([&](int j) { return Lk[j]; }, n, kk, hist, sh_sel);
  // survivors: upper bound reaches the k-th best lower bound.  Lk is dead now -> survivor list.
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    if (Uk[j] >= v) {
      const int s = atomicAdd(&sh_ns, 1);
      Lk[s] = static_cast<uint32_t>(j);
    }
  }
  __syncthreads();
  const int ns = sh_ns;
  if (ns > static_cast<int>(cand_cap / 2)) {  // (uniform) the survivors' keys would not fit: massive ties
    if (threadIdx.x == 0) overflow_flags[qq] = 1u;
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* b = q + static_cast<size_t>(qq) * d;
  for (int s = warp; s < ns; s += nw) {
    const uint32_t i = Ix[Lk[s]];
    const double dot = warp_dot(feat + static_cast<int64_t>(i) * d, b, d, lane);
    if (lane == 0) {
      const float inv = normalize ? inv_of_norm(row_norm[i]) : 1.f;
      const float f = canon_score(dot, inv, normalize, scale, qq);
      K64[s] = (static_cast<unsigned long long>(f2ord(f)) << 32) | (0xFFFFFFFFu - i);
    }
  }
  __syncthreads();
  const int kf = min(k, ns);
  unsigned long long v64 = 0ull;  // keys are unique (row id in the low word): rank by counting
  if (kf > 0 && ns > 2048) v64 = block_kth_largest<unsigned long long>([&](int j) { return K64[j]; }, ns, kf, &sh_cnt);
  for (int s = threadIdx.x; s < ns; s += blockDim.x) {
    const unsigned long long key = K64[s];
    if (kf > 0 && key >= v64) {
      int rank = 0;
      for (int t = 0; t < ns; ++t) rank += (K64[t] > key) ? 1 : 0;
      if (rank < kf) {
        out_idx[static_cast<size_t>(qq) * k + rank] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
        out_score[static_cast<size_t>(qq) * k + rank] = ord2f(static_cast<uint32_t>(key >> 32));
      }
    }
  }
  for (int s = kf + threadIdx.x; s < k; s += blockDim.x) {
    out_idx[static_cast<size_t>(qq) * k + s] = -1;
    out_score[static_cast<size_t>(qq) * k + s] = -INFINITY;
  }
}

// ------------------------------------------------------------------ exact top-k of a vector
// level 1: each block selects the k best of its chunk (keys in shared memory);
// level 2: one block selects the k best of the survivors and sorts them.
constexpr int kVecChunk = 4096;

__device__ __forceinline__ unsigned long long vec_key(float f, int64_t i) {
  return (static_cast<unsigned long long>(f2ord(f)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(i));
}

__global__ void __launch_bounds__(256)
topk_vec_l1_kernel(const float* __restrict__ v, int64_t n, int32_t k, unsigned long long* __restrict__ out) {
  __shared__ unsigned long long keys[kVecChunk];
  __shared__ int sh_cnt;
  __shared__ int sh_w;
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kVecChunk;
  const int m = static_cast<int>(min(static_cast<int64_t>(kVecChunk), n - base));
  for (int j = threadIdx.x; j < m; j += blockDim.x) keys[j] = vec_key(v[base + j], base + j);
  if (threadIdx.x == 0) sh_w = 0;
  __syncthreads();
  const int kk = min(k, m);
  const unsigned long long t =
      block_kth_largest<unsigned long long>([&](int j) { return keys[j]; }, m, kk, &sh_cnt);
  unsigned long long* o = out + static_cast<size_t>(blockIdx.x) * k;
  for (int j = threadIdx.x; j < m; j += blockDim.x)
    if (keys[j] >= t) o[atomicAdd(&sh_w, 1)] = keys[j];
  __syncthreads();
  for (int j = sh_w + threadIdx.x; j < k; j += blockDim.x) o[j] = 0ull;  // 0 = "nothing"
}

__global__ void __launch_bounds__(1024)
topk_vec_l2_kernel(const unsigned long long* __restrict__ keys, int m, int32_t k, int64_t n,
                   int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  __shared__ int sh_cnt;
  __shared__ unsigned long long sel[AVL_MAX_TOPK];
  __shared__ int sh_w;
  if (threadIdx.x == 0) sh_w = 0;
  __syncthreads();
  const int kk = static_cast<int>(min(static_cast<int64_t>(k), n));
  const unsigned long long t =
      block_kth_largest<unsigned long long>([&](int j) { return keys[j]; }, m, kk, &sh_cnt);
  for (int j = threadIdx.x; j < m; j += blockDim.x) {
    const unsigned long long key = keys[j];
    if (key >= t && key != 0ull) {
      const int s = atomicAdd(&sh_w, 1);
      if (s < AVL_MAX_TOPK) sel[s] = key;
    }
  }
  __syncthreads();
  const int ns = min(sh_w, kk);
  for (int s = threadIdx.x; s < k; s += blockDim.x) {
    if (s < ns) {
      const unsigned long long key = sel[s];
      int rank = 0;
      for (int u = 0; u < ns; ++u) rank += (sel[u] > key) ? 1 : 0;
      out_idx[rank] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
      out_val[rank] = ord2f(static_cast<uint32_t>(key >> 32));
    }
  }
  __syncthreads();
  for (int s = ns + threadIdx.x; s < k; s += blockDim.x) {
    out_idx[s] = -1;
    out_val[s] = -INFINITY;
  }
}

// ------------------------------------------------------------------ exact fallback, decided on the device
// Always launched after topk_finalize; every block walks the overflow flags and leaves at once when none is set (the
// normal case: one short launch, no host round trip -- the call stays asynchronous).  For a flagged query (its
// candidate buckets overflowed: adversarial data, massive ties) the whole column is re-scored exactly: every warp
// keeps the top-k keys of its rows in shared memory, the block merges its warps' lists, the LAST block to finish
// (ticket) merges the blocks' lists.  keys = (ordered score bits << 32) | ~row: unique, so (score desc, row asc).
constexpr int kFbThreads = 256;
constexpr int kFbGroup = 8;   // flagged queries scored per pass over the map
__global__ void __launch_bounds__(kFbThreads)
topk_fallback_kernel(const float* __restrict__ feat, int64_t n, int32_t d, const float* __restrict__ q, int32_t nq,
                     const float* __restrict__ scale, const float* __restrict__ row_norm, int normalize, int32_t k,
                     int32_t group, const uint32_t* __restrict__ overflow_flags, unsigned long long* __restrict__ scratch,
                     uint32_t* __restrict__ tickets, int64_t* __restrict__ out_idx, float* __restrict__ out_score) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int kWarps = kFbThreads / 32;
  extern __shared__ uint8_t fb_sm[];
  // dynamic: [group][d] doubles of the group's queries (converted once: float -> double conversions run at a quarter
  // of the fp64 FMA rate and would otherwise be two per product) | [kWarps][group][k] keys (per-warp sorted lists)
  double* qs = reinterpret_cast<double*>(fb_sm);
  unsigned long long* wl = reinterpret_cast<unsigned long long*>(fb_sm + static_cast<size_t>(group) * d * 8);
  __shared__ int wcnt[kWarps][kFbGroup];
  __shared__ int sh_flagged[AVL_MAX_QUERIES];
  __shared__ int sh_nflag;
  __shared__ int sh_cnt;
  __shared__ uint32_t sh_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t gw = static_cast<int64_t>(blockIdx.x) * kWarps + warp, nwarps = static_cast<int64_t>(gridDim.x) * kWarps;
  // the normal case first: one flag per thread, one barrier, gone (walking the 256 flags one dependent load after the
  // other cost 24 us per call on B200)
  static_assert(kFbThreads >= AVL_MAX_QUERIES, "one flag per thread");
  if (!__syncthreads_or(static_cast<int>(threadIdx.x) < nq && overflow_flags[threadIdx.x] != 0u)) return;
  if (threadIdx.x == 0) {  // the flagged queries in ascending order (the flags are read-only during this launch)
    int c = 0;
    for (int i = 0; i < nq; ++i)
      if (overflow_flags[i]) sh_flagged[c++] = i;
    sh_nflag = c;
  }
  __syncthreads();
  const int nflag = sh_nflag;
  // Up to `group` flagged queries share ONE pass over the map: a row is loaded once (512 elements at a time, 16 per
  // lane, in registers) and multiplied into every query of the group, so 256 flagged queries cost 32 passes, not 256.
  for (int g0 = 0; g0 < nflag; g0 += group) {
    const int ng = min(group, nflag - g0);
    for (int e = threadIdx.x; e < ng * d; e += blockDim.x) {
      const int g = e / d, c = e - g * d;
      qs[e] = static_cast<double>(q[static_cast<size_t>(sh_flagged[g0 + g]) * d + c]);
    }
    __syncthreads();
    // ---- phase 1: per-warp top-k of its rows for every query of the group (sorted descending, insertion by lane 0)
    int cnt[kFbGroup];
    unsigned long long kmin[kFbGroup];
#pragma unroll
    for (int g = 0; g < kFbGroup; ++g) { cnt[g] = 0; kmin[g] = 0ull; }
    for (int64_t row = gw; row < n; row += nwarps) {
      double acc[kFbGroup];
#pragma unroll
      for (int g = 0; g < kFbGroup; ++g) acc[g] = 0.0;
      const float* a = feat + row * d;
      for (int k0 = 0; k0 < d; k0 += 512) {   // same per-lane order as warp_dot: k = lane, lane + 32, ... ascending
        float af[16];
        double av[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const int c = k0 + t * 32 + lane;
          af[t] = c < d ? __ldg(a + c) : 0.f;
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) av[t] = static_cast<double>(af[t]);
#pragma unroll
        for (int g = 0; g < kFbGroup; ++g) {
          if (g < ng) {
            const double* qg = qs + g * d;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
              const int c = k0 + t * 32 + lane;
              if (c < d) acc[g] = fma(av[t], qg[c], acc[g]);
            }
          }
        }
      }
      const float inv = normalize ? inv_of_norm(row_norm[row]) : 1.f;
#pragma unroll
      for (int g = 0; g < kFbGroup; ++g) {
        if (g < ng) {
          const double dot = warp_sum(acc[g]);
          if (lane == 0) {
            unsigned long long* list = wl + (static_cast<size_t>(warp) * group + g) * k;
            const unsigned long long key = vec_key(canon_score(dot, inv, normalize, scale, sh_flagged[g0 + g]), row);
            if (cnt[g] < k || key > kmin[g]) {
              int pos = cnt[g] < k ? cnt[g] : k - 1;
              while (pos > 0 && list[pos - 1] < key) { list[pos] = list[pos - 1]; --pos; }
              list[pos] = key;
              if (cnt[g] < k) ++cnt[g];
              if (cnt[g] == k) kmin[g] = list[k - 1];
            }
          }
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < kFbGroup; ++g) wcnt[warp][g] = cnt[g];
    }
    __syncthreads();
    for (int g = 0; g < ng; ++g) {
      const int qq = sh_flagged[g0 + g];
      // ---- phase 2: block merge by rank counting (<= 8 * 128 keys), block list -> scratch[qq][block][k] (0 = nothing)
      unsigned long long* mine = scratch + (static_cast<size_t>(qq) * gridDim.x + blockIdx.x) * k;
      for (int j = threadIdx.x; j < k; j += blockDim.x) mine[j] = 0ull;
      __syncthreads();
      for (int e = threadIdx.x; e < kWarps * k; e += blockDim.x) {
        const int w = e / k, j = e - w * k;
        if (j >= wcnt[w][g]) continue;
        const unsigned long long key = wl[(static_cast<size_t>(w) * group + g) * k + j];
        int rank = 0;
        for (int w2 = 0; w2 < kWarps; ++w2)
          for (int j2 = 0; j2 < wcnt[w2][g]; ++j2) rank += (wl[(static_cast<size_t>(w2) * group + g) * k + j2] > key) ? 1 : 0;
        if (rank < k) mine[rank] = key;
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) sh_last = (atomicAdd(tickets + qq, 1u) == gridDim.x - 1) ? 1u : 0u;
      __syncthreads();
      if (sh_last) {
        // ---- phase 3 (last block): top-k of the gridDim.x * k block keys
        __threadfence();
        const unsigned long long* all = scratch + static_cast<size_t>(qq) * gridDim.x * k;
        const int m = static_cast<int>(gridDim.x) * k;
        const int kk = static_cast<int>(min(static_cast<int64_t>(k), n));
        const unsigned long long t =
            block_kth_largest<unsigned long long>([&](int j) { return __ldcg(all + j); }, m, kk, &sh_cnt);
        for (int s = threadIdx.x; s < k; s += blockDim.x) {
          out_idx[static_cast<size_t>(qq) * k + s] = -1;
          out_score[static_cast<size_t>(qq) * k + s] = -INFINITY;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) {
          const unsigned long long key = __ldcg(all + j);
          if (key == 0ull || key < t) continue;
          int rank = 0;
          for (int u = 0; u < m; ++u) {
            const unsigned long long o = __ldcg(all + u);
            rank += (o > key) ? 1 : 0;
          }
          if (rank < kk) {
            out_idx[static_cast<size_t>(qq) * k + rank] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
            out_score[static_cast<size_t>(qq) * k + rank] = ord2f(static_cast<uint32_t>(key >> 32));
          }
        }
        if (threadIdx.x == 0) tickets[qq] = 0u;  // ready for the next call
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ merge of per-shard top-k
// idx/val: (S, Q, k) per-shard results with GLOBAL row ids (-1 = empty slot).  One block per query:
// rank every entry by (score desc, row asc) by counting -- S*k <= 1024 entries.
__global__ void __launch_bounds__(256)
merge_topk_kernel(const int64_t* __restrict__ idx, const float* __restrict__ val, int32_t n_shards, int32_t nq,
                  int32_t k, int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  __shared__ int64_t si[1024];
  __shared__ float sv[1024];
  const int q = blockIdx.x;
  const int m = n_shards * k;
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    const int s = e / k, j = e - s * k;
    const size_t off = (static_cast<size_t>(s) * nq + q) * k + j;
    si[e] = idx[off];
    sv[e] = val[off];
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    out_idx[static_cast<size_t>(q) * k + j] = -1;
    out_val[static_cast<size_t>(q) * k + j] = -INFINITY;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m; e += blockDim.x) {
    const int64_t i = si[e];
    if (i < 0) continue;
    const float v = sv[e];
    int rank = 0;
    for (int t = 0; t < m; ++t) {
      const int64_t it = si[t];
      const float vt = sv[t];
      rank += (it >= 0 && (vt > v || (vt == v && it < i))) ? 1 : 0;
    }
    if (rank < k) {
      out_idx[static_cast<size_t>(q) * k + rank] = i;
      out_val[static_cast<size_t>(q) * k + rank] = v;
    }
  }
}

// ------------------------------------------------------------------ fusion helpers
// per-column min / max of a column-major matrix m[col][row] (each column contiguous).
__global__ void __launch_bounds__(256)
minmax_cols_kernel(const float* __restrict__ m, int64_t n, uint32_t* __restrict__ out_min,
                   uint32_t* __restrict__ out_max) {
  const int col = blockIdx.y;
  const float* c = m + static_cast<int64_t>(col) * n;
  uint32_t lo = 0xFFFFFFFFu, hi = 0u;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t key = f2ord(c[i]);
    lo = min(lo, key);
    hi = max(hi, key);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(out_min + col, lo);
    atomicMax(out_max + col, hi);
  }
}

__device__ __forceinline__ float minmax_norm(float s, float lo, float hi) {
  // (scores - min) / (max - min), fp32 like the reference's numpy float32 arithmetic
  return __fdiv_rn(__fsub_rn(s, lo), __fsub_rn(hi, lo));
}

__global__ void __launch_bounds__(256)
fuse_heat_kernel(const float* __restrict__ sa, const float* __restrict__ sb, int64_t n, int32_t pair,
                 const uint32_t* __restrict__ min_a, const uint32_t* __restrict__ max_a,
                 const uint32_t* __restrict__ min_b, const uint32_t* __restrict__ max_b, int32_t combine,
                 float* __restrict__ heat) {
  const float la = ord2f(min_a[pair]), ha = ord2f(max_a[pair]);
  const float lb = ord2f(min_b[pair]), hb = ord2f(max_b[pair]);
  const float* ca = sa + static_cast<int64_t>(pair) * n;
  const float* cb = sb + static_cast<int64_t>(pair) * n;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = minmax_norm(ca[i], la, ha);
    const float y = minmax_norm(cb[i], lb, hb);
    float h;
    if (combine == AVL_FUSE_PRODUCT) h = __fmul_rn(x, y);
    else if (combine == AVL_FUSE_MAX) h = fmaxf(x, y);
    else h = __fadd_rn(x, y);
    heat[i] = h;
  }
}

// =================================================================== cross-modal fusion through the screen
// avl_fuse_topk without exact dense columns (BASELINE config 3).  Both modalities are screened by the
// tcgen05 kernel in dense mode -> s~ (pairs, n) column-major, with the rigorous band
//     |s~_ij - dot_ij| <= r_i * bn_j,   r_i = (rho * ||a~_i|| + c_i) * (1 + 1e-6)       
// Everything the reference computes from the exact scores is monotone in them -- the column min / max
// (sound_map.py:151-152), (s - min) / (max - min) in fp32, product | max | sum of two values in [0, 1] -- so
// interval bounds propagate through the SAME fp32 operation sequence and decide which rows must be re-scored
// exactly (fp64-accumulated dots, the canonical score); the returned ids / heats equal the exact path's.
struct FuseSide {
  const float* dense;     // (pairs, n) column-major screen scores
  const float* row_c;
  const float* row_an;
  const float* row_norm;
  const float* q_bn;
  const float* q_glob;    // [0] = rho
  const float* scale;     // per pair or null
  const float* feat;      // fp32 rows for the exact re-score
  const float* q;         // (pairs, d) fp32 queries
  int32_t d;
  int32_t normalize;
};

struct RowBand { float r, invw; };

__device__ __forceinline__ RowBand fuse_row_band(const FuseSide& m, int64_t row) {
  RowBand b;
  b.r = fmaf(__ldg(m.q_glob), __ldg(m.row_an + row), __ldg(m.row_c + row)) * 1.000001f;
  b.invw = m.normalize ? 1.f / fmaxf(__ldg(m.row_norm + row), 1e-30f) : 1.f;
  return b;
}
// bounds of the canonical score fl(fl(fl(dot) * inv) * scale): the interval of the real-number value, widened by a
// relative 4e-6 (>> the 3 fp32 roundings of the canonical sequence and the ones made here)
__device__ __forceinline__ void fuse_bounds(const FuseSide& m, const RowBand& b, float s, int j, float& lo, float& hi) {
  const float e = b.r * __ldg(m.q_bn + j);
  lo = (s - e) * b.invw;
  hi = (s + e) * b.invw;
  if (m.scale) {
    const float sc = __ldg(m.scale + j);  // > 0 (checked by the caller)
    lo *= sc;
    hi *= sc;
  }
  lo = fmaf(-fabsf(lo), 4e-6f, lo);
  hi = fmaf(fabsf(hi), 4e-6f, hi);
}
__device__ __forceinline__ float fuse_combine(float x, float y, int combine) {
  if (combine == AVL_FUSE_PRODUCT) return __fmul_rn(x, y);
  if (combine == AVL_FUSE_MAX) return fmaxf(x, y);
  return __fadd_rn(x, y);
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

constexpr int kFuseChunk = 8;   // columns per block in the column passes

// V consecutive rows per thread: V = 4 turns every load of the column passes into a 16-byte one (n % 4 == 0 keeps
// the column bases aligned), which is what lets them stream near HBM speed; V = 1 is the general fallback.
template <int V> struct FVec { float v[V]; };
template <int V>
__device__ __forceinline__ FVec<V> fuse_ld(const float* __restrict__ p) {
  FVec<V> r;
  if constexpr (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}
template <int V> struct RowBandV { float r[V], invw[V]; };
template <int V>
__device__ __forceinline__ RowBandV<V> fuse_row_band_v(const FuseSide& m, int64_t row) {
  RowBandV<V> b;
  const float rho = __ldg(m.q_glob);
  const FVec<V> an = fuse_ld<V>(m.row_an + row), c = fuse_ld<V>(m.row_c + row);
  FVec<V> nr;
  if (m.normalize) nr = fuse_ld<V>(m.row_norm + row);
#pragma unroll
  for (int t = 0; t < V; ++t) {
    b.r[t] = fmaf(rho, an.v[t], c.v[t]) * 1.000001f;
    b.invw[t] = m.normalize ? 1.f / fmaxf(nr.v[t], 1e-30f) : 1.f;
  }
  return b;
}
template <int V>
__device__ __forceinline__ RowBand band_of(const RowBandV<V>& b, int t) {
  RowBand o;
  o.r = b.r[t];
  o.invw = b.invw[t];
  return o;
}
__device__ __forceinline__ void fuse_bounds_q(const FuseSide& m, const RowBand& b, float s, float bn, float sc,
                                              float& lo, float& hi) {  // fuse_bounds with the per-query values hoisted
  const float e = b.r * bn;
  lo = (s - e) * b.invw * sc;
  hi = (s + e) * b.invw * sc;
  lo = fmaf(-fabsf(lo), 4e-6f, lo);
  hi = fmaf(fabsf(hi), 4e-6f, hi);
}

// pass 1: per column max of the lower bounds and min of the upper bounds (ordered-uint atomics)
template <int V>
__global__ void __launch_bounds__(256, 2)
fuse_colstats_kernel(const FuseSide sa, const FuseSide sb, int64_t n, int32_t pairs, uint32_t* __restrict__ max_lb,
                     uint32_t* __restrict__ min_ub) {
  const FuseSide& m = blockIdx.z ? sb : sa;
  const int j0 = blockIdx.y * kFuseChunk;
  uint32_t mx[kFuseChunk], mn[kFuseChunk];
#pragma unroll
  for (int c = 0; c < kFuseChunk; ++c) { mx[c] = 0u; mn[c] = 0xFFFFFFFFu; }
  for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * V; row < n;
       row += static_cast<int64_t>(gridDim.x) * blockDim.x * V) {
    const RowBandV<V> b = fuse_row_band_v<V>(m, row);
    FVec<V> sv[kFuseChunk];
#pragma unroll
    for (int c = 0; c < kFuseChunk; ++c)
      if (j0 + c < pairs) sv[c] = fuse_ld<V>(m.dense + static_cast<int64_t>(j0 + c) * n + row);
#pragma unroll
    for (int c = 0; c < kFuseChunk; ++c) {
      const int j = j0 + c;
      if (j < pairs) {
        const float bn = __ldg(m.q_bn + j), sc = m.scale ? __ldg(m.scale + j) : 1.f;
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float lo, hi;
          fuse_bounds_q(m, band_of<V>(b, t), sv[c].v[t], bn, sc, lo, hi);
          mx[c] = max(mx[c], f2ord(lo));
          mn[c] = min(mn[c], f2ord(hi));
        }
      }
    }
  }
  __shared__ uint32_t smx[kFuseChunk], smn[kFuseChunk];
  if (threadIdx.x < kFuseChunk) { smx[threadIdx.x] = 0u; smn[threadIdx.x] = 0xFFFFFFFFu; }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < kFuseChunk; ++c) {
    const uint32_t a = __reduce_max_sync(0xffffffffu, mx[c]);
    const uint32_t i = __reduce_min_sync(0xffffffffu, mn[c]);
    if ((threadIdx.x & 31) == 0) { atomicMax(smx + c, a); atomicMin(smn + c, i); }
  }
  __syncthreads();
  if (threadIdx.x < kFuseChunk && j0 + threadIdx.x < pairs) {
    const int o = blockIdx.z * pairs + j0 + threadIdx.x;
    atomicMax(max_lb + o, smx[threadIdx.x]);
    atomicMin(min_ub + o, smn[threadIdx.x]);
  }
}

// pass 2: rows that can hold a column's exact max (ub >= max lb) or min (lb <= min ub)
template <int V>
__global__ void __launch_bounds__(256, 2)
fuse_collect_extreme_kernel(const FuseSide sa, const FuseSide sb, int64_t n, int32_t pairs,
                            const uint32_t* __restrict__ max_lb, const uint32_t* __restrict__ min_ub,
                            uint32_t* __restrict__ ext_cnt, uint32_t* __restrict__ ext_row, uint32_t ext_cap) {
  const FuseSide& m = blockIdx.z ? sb : sa;
  const int j0 = blockIdx.y * kFuseChunk;
  for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * V; row < n;
       row += static_cast<int64_t>(gridDim.x) * blockDim.x * V) {
    const RowBandV<V> b = fuse_row_band_v<V>(m, row);
    FVec<V> sv[kFuseChunk];
#pragma unroll
    for (int c = 0; c < kFuseChunk; ++c)
      if (j0 + c < pairs) sv[c] = fuse_ld<V>(m.dense + static_cast<int64_t>(j0 + c) * n + row);
#pragma unroll
    for (int c = 0; c < kFuseChunk; ++c) {
      const int j = j0 + c;
      if (j < pairs) {
        const int o = blockIdx.z * pairs + j;
        const float bn = __ldg(m.q_bn + j), sc = m.scale ? __ldg(m.scale + j) : 1.f;
        const uint32_t tmax = __ldg(max_lb + o), tmin = __ldg(min_ub + o);
#pragma unroll
        for (int t = 0; t < V; ++t) {
          float lo, hi;
          fuse_bounds_q(m, band_of<V>(b, t), sv[c].v[t], bn, sc, lo, hi);
          if (f2ord(hi) >= tmax) {
            const uint32_t pos = atomicAdd(ext_cnt + 2 * o, 1u);
            if (pos < ext_cap) ext_row[static_cast<size_t>(2 * o) * ext_cap + pos] = static_cast<uint32_t>(row + t);
          }
          if (f2ord(lo) <= tmin) {
            const uint32_t pos = atomicAdd(ext_cnt + 2 * o + 1, 1u);
            if (pos < ext_cap) ext_row[static_cast<size_t>(2 * o + 1) * ext_cap + pos] = static_cast<uint32_t>(row + t);
          }
        }
      }
    }
  }
}

// pass 3: exact column max / min from the candidates.  grid (pairs, 2 sides), one warp per candidate.
__global__ void __launch_bounds__(256)
fuse_exact_extreme_kernel(const FuseSide sa, const FuseSide sb, int32_t pairs, const uint32_t* __restrict__ ext_cnt,
                          const uint32_t* __restrict__ ext_row, uint32_t ext_cap, float* __restrict__ mm,
                          uint32_t* __restrict__ overflow) {
  const FuseSide& m = blockIdx.y ? sb : sa;
  const int j = blockIdx.x, o = blockIdx.y * pairs + j;
  __shared__ uint32_t s_ext[2];
  if (threadIdx.x == 0) { s_ext[0] = 0u; s_ext[1] = 0xFFFFFFFFu; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* b = m.q + static_cast<size_t>(j) * m.d;
  for (int kind = 0; kind < 2; ++kind) {
    const uint32_t cnt = ext_cnt[2 * o + kind];
    if (cnt > ext_cap) {
      if (threadIdx.x == 0) atomicExch(overflow, 1u);
      continue;
    }
    const uint32_t* rows = ext_row + static_cast<size_t>(2 * o + kind) * ext_cap;
    for (uint32_t c = warp; c < cnt; c += nw) {
      const uint32_t i = rows[c];
      const double dot = warp_dot(m.feat + static_cast<int64_t>(i) * m.d, b, m.d, lane);
      if (lane == 0) {
        const float inv = m.normalize ? inv_of_norm(m.row_norm[i]) : 1.f;
        const uint32_t key = f2ord(canon_score(dot, inv, m.normalize, m.scale, j));
        if (kind == 0) atomicMax(s_ext + 0, key); else atomicMin(s_ext + 1, key);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mm[(2 * blockIdx.y + 0) * pairs + j] = ord2f(s_ext[1]);  // exact min
    mm[(2 * blockIdx.y + 1) * pairs + j] = ord2f(s_ext[0]);  // exact max
  }
}

// heat bound of (row, pair): the fp32 sequence of fuse_heat_kernel applied to the clamped score bounds
template <bool kUpper>
__device__ __forceinline__ float fuse_heat_bound(const FuseSide& sa, const FuseSide& sb, const RowBand& ba,
                                                 const RowBand& bb, int64_t n, int64_t row, int j, int pairs,
                                                 const float* __restrict__ mm, int combine) {
  float la, ha, lb, hb;
  fuse_bounds(sa, ba, __ldg(sa.dense + static_cast<int64_t>(j) * n + row), j, la, ha);
  fuse_bounds(sb, bb, __ldg(sb.dense + static_cast<int64_t>(j) * n + row), j, lb, hb);
  const float mna = __ldg(mm + j), mxa = __ldg(mm + pairs + j), mnb = __ldg(mm + 2 * pairs + j), mxb = __ldg(mm + 3 * pairs + j);
  const float x = minmax_norm(clampf(kUpper ? ha : la, mna, mxa), mna, mxa);
  const float y = minmax_norm(clampf(kUpper ? hb : lb, mnb, mxb), mnb, mxb);
  return fuse_combine(x, y, combine);
}

// pass 4: lower bounds of the heat on a strided row sample -> (pairs, n_sample) for select_threshold
__global__ void __launch_bounds__(256)
fuse_sample_kernel(const FuseSide sa, const FuseSide sb, int64_t n, int32_t pairs, int64_t stride, int32_t n_sample,
                   const float* __restrict__ mm, int combine, float* __restrict__ sample_t) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_sample) return;
  const int64_t row = static_cast<int64_t>(t) * stride;
  if (row >= n) {
    for (int j = 0; j < pairs; ++j) sample_t[static_cast<int64_t>(j) * n_sample + t] = -INFINITY;
    return;
  }
  const RowBand ba = fuse_row_band(sa, row), bb = fuse_row_band(sb, row);
  for (int j = 0; j < pairs; ++j) {
    const float h = fuse_heat_bound<false>(sa, sb, ba, bb, n, row, j, pairs, mm, combine);
    sample_t[static_cast<int64_t>(j) * n_sample + t] = h == h ? h : -INFINITY;  // NaN (max == min): no threshold
  }
}

// heat bounds of one (row, pair) from the two screen scores
__device__ __forceinline__ void fuse_heat_bounds2(const FuseSide& sa, const FuseSide& sb, const RowBand& ba, const RowBand& bb,
                                                  float s_a, float s_b, float bna, float sca, float bnb, float scb,
                                                  float mna, float mxa, float mnb, float mxb, int combine, float& h_lo,
                                                  float& h_hi) {
  float la, ha, lb, hb;
  fuse_bounds_q(sa, ba, s_a, bna, sca, la, ha);
  fuse_bounds_q(sb, bb, s_b, bnb, scb, lb, hb);
  h_hi = fuse_combine(minmax_norm(clampf(ha, mna, mxa), mna, mxa), minmax_norm(clampf(hb, mnb, mxb), mnb, mxb), combine);
  h_lo = fuse_combine(minmax_norm(clampf(la, mna, mxa), mna, mxa), minmax_norm(clampf(lb, mnb, mxb), mnb, mxb), combine);
}

// pass 5: rows whose heat upper bound reaches the pair's threshold, with both heat bounds
template <int V>
__global__ void __launch_bounds__(256, 3)
fuse_collect_heat_kernel(const FuseSide sa, const FuseSide sb, int64_t n, int32_t pairs, const float* __restrict__ mm,
                         int combine, const float* __restrict__ thr, uint32_t* __restrict__ cand_cnt,
                         uint32_t* __restrict__ cand_row, float* __restrict__ cand_lo, float* __restrict__ cand_hi,
                         uint32_t cand_cap) {
  constexpr int kC = kFuseChunk / 2;  // two modalities per pair: half the columns per step
  const int j0 = blockIdx.y * kFuseChunk;
  for (int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * V; row < n;
       row += static_cast<int64_t>(gridDim.x) * blockDim.x * V) {
    const RowBandV<V> ba = fuse_row_band_v<V>(sa, row), bb = fuse_row_band_v<V>(sb, row);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      FVec<V> va[kC], vb[kC];
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const int j = j0 + h * kC + c;
        if (j < pairs) {
          va[c] = fuse_ld<V>(sa.dense + static_cast<int64_t>(j) * n + row);
          vb[c] = fuse_ld<V>(sb.dense + static_cast<int64_t>(j) * n + row);
        }
      }
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        const int j = j0 + h * kC + c;
        if (j >= pairs) continue;
        const float bna = __ldg(sa.q_bn + j), sca = sa.scale ? __ldg(sa.scale + j) : 1.f;
        const float bnb = __ldg(sb.q_bn + j), scb = sb.scale ? __ldg(sb.scale + j) : 1.f;
        const float mna = __ldg(mm + j), mxa = __ldg(mm + pairs + j), mnb = __ldg(mm + 2 * pairs + j), mxb = __ldg(mm + 3 * pairs + j);
        const float tj = __ldg(thr + j);
        // division-free prefilter: (x - min) * (1 / range, rounded up) * (1 + 2^-21) >= fl((x - min) / range), so
        // the cheap product bounds the real upper bound from above; the IEEE divisions (4 per element, which made
        // this pass compute-bound) run only for the few rows that pass
        const float ira = __fdiv_ru(1.0000005f, __fsub_rd(mxa, mna)), irb = __fdiv_ru(1.0000005f, __fsub_rd(mxb, mnb));
#pragma unroll
        for (int t = 0; t < V; ++t) {
          {
            const RowBand ra = band_of<V>(ba, t), rb = band_of<V>(bb, t);
            float la, ha, lb, hb;
            fuse_bounds_q(sa, ra, va[c].v[t], bna, sca, la, ha);
            fuse_bounds_q(sb, rb, vb[c].v[t], bnb, scb, lb, hb);
            const float xu = fminf(__fmul_ru(__fsub_ru(clampf(ha, mna, mxa), mna), ira), 1.f);
            const float yu = fminf(__fmul_ru(__fsub_ru(clampf(hb, mnb, mxb), mnb), irb), 1.f);
            const float hu = combine == AVL_FUSE_PRODUCT ? __fmul_ru(xu, yu) : (combine == AVL_FUSE_MAX ? fmaxf(xu, yu) : __fadd_ru(xu, yu));
            if (hu < tj) continue;  // NaN (degenerate column) falls through to the exact-sequence test below
          }
          float h_lo, h_hi;
          fuse_heat_bounds2(sa, sb, band_of<V>(ba, t), band_of<V>(bb, t), va[c].v[t], vb[c].v[t], bna, sca, bnb, scb, mna, mxa,
                            mnb, mxb, combine, h_lo, h_hi);
          if (!(h_hi < tj)) {  // also true for NaN: degenerate columns go to the exact path via overflow
            const uint32_t pos = atomicAdd(cand_cnt + j, 1u);
            if (pos < cand_cap) {
              const size_t o = static_cast<size_t>(j) * cand_cap + pos;
              cand_row[o] = static_cast<uint32_t>(row + t);
              cand_hi[o] = h_hi;
              cand_lo[o] = h_lo;
            }
          }
        }
      }
    }
  }
}

// pass 6, one block per pair: k-th best LOWER bound over the candidates -> survivors (upper bound reaches it) ->
// exact heat of the survivors (fp64-accumulated dots of both modalities) -> top-k by (heat desc, row asc).
constexpr int kFuseSurvivors = 2048;
__global__ void __launch_bounds__(1024)
fuse_finalize_kernel(const FuseSide sa, const FuseSide sb, int32_t pairs, const float* __restrict__ mm, int combine,
                     int32_t k, const uint32_t* __restrict__ cand_cnt, const uint32_t* __restrict__ cand_row,
                     const float* __restrict__ cand_lo, const float* __restrict__ cand_hi, uint32_t cand_cap,
                     int64_t* __restrict__ out_idx, float* __restrict__ out_heat, uint32_t* __restrict__ overflow) {
  extern __shared__ uint8_t sm[];
  uint32_t* Lk = reinterpret_cast<uint32_t*>(sm);   // [cand_cap] ordered lower bounds
  __shared__ unsigned long long K64[kFuseSurvivors];
  __shared__ uint32_t S[kFuseSurvivors];            // candidate slots of the survivors
  __shared__ int sh_cnt;
  __shared__ int sh_ns;
  const int j = blockIdx.x;
  const uint32_t cnt = cand_cnt[j];
  if (cnt > cand_cap) {
    if (threadIdx.x == 0) atomicExch(overflow, 1u);
    return;
  }
  const int n = static_cast<int>(cnt);
  const size_t base = static_cast<size_t>(j) * cand_cap;
  if (threadIdx.x == 0) sh_ns = 0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) Lk[c] = max(f2ord(cand_lo[base + c]), 1u);
  __syncthreads();
  const int kk = min(k, n);
  uint32_t v = 0u;
  __shared__ uint32_t hist[256];
  __shared__ int sh_sel[2];
  if (kk > 0) v = block_kth_largest_radix32([&](int c) { return Lk[c]; }, n, kk, hist, sh_sel);
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    if (max(f2ord(cand_hi[base + c]), 1u) >= v) {
      const int t = atomicAdd(&sh_ns, 1);
      if (t < kFuseSurvivors) S[t] = static_cast<uint32_t>(c);
    }
  }
  __syncthreads();
  const int ns = sh_ns;
  if (ns > kFuseSurvivors) {  // massive ties: the exact path answers the call
    if (threadIdx.x == 0) atomicExch(overflow, 1u);
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float mna = mm[j], mxa = mm[pairs + j], mnb = mm[2 * pairs + j], mxb = mm[3 * pairs + j];
  const float* qa = sa.q + static_cast<size_t>(j) * sa.d;
  const float* qb = sb.q + static_cast<size_t>(j) * sb.d;
  for (int c = warp; c < ns; c += nw) {
    const uint32_t i = cand_row[base + S[c]];
    const double da = warp_dot(sa.feat + static_cast<int64_t>(i) * sa.d, qa, sa.d, lane);
    const double db = warp_dot(sb.feat + static_cast<int64_t>(i) * sb.d, qb, sb.d, lane);
    if (lane == 0) {
      const float fa = canon_score(da, sa.normalize ? inv_of_norm(sa.row_norm[i]) : 1.f, sa.normalize, sa.scale, j);
      const float fb = canon_score(db, sb.normalize ? inv_of_norm(sb.row_norm[i]) : 1.f, sb.normalize, sb.scale, j);
      const float h = fuse_combine(minmax_norm(fa, mna, mxa), minmax_norm(fb, mnb, mxb), combine);
      K64[c] = (static_cast<unsigned long long>(f2ord(h)) << 32) | (0xFFFFFFFFu - i);
    }
  }
  __syncthreads();
  const int kf = min(k, ns);
  for (int c = threadIdx.x; c < ns; c += blockDim.x) {
    const unsigned long long key = K64[c];
    int rank = 0;
    for (int t = 0; t < ns; ++t) rank += (K64[t] > key) ? 1 : 0;
    if (rank < kf) {
      out_idx[static_cast<size_t>(j) * k + rank] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(key));
      out_heat[static_cast<size_t>(j) * k + rank] = ord2f(static_cast<uint32_t>(key >> 32));
    }
  }
  for (int c = kf + threadIdx.x; c < k; c += blockDim.x) {
    out_idx[static_cast<size_t>(j) * k + c] = -1;
    out_heat[static_cast<size_t>(j) * k + c] = -INFINITY;
  }
}

__global__ void fill_u32_kernel(uint32_t* p, int n, uint32_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

// =================================================================== launchers
// Kernels that may share an SM with a CTA of the screen kernel (217 KiB of shared memory) ask for the maximum
// shared-memory carveout too: an SM cannot change its L1 / shared split while blocks are resident, so a small kernel
// that leaves it at the default split keeps the screen CTA out until it has drained (seen as a 2x slower head of the next
// call when the tail of a pipelined call ran on the default carveout).
template <typename K>
static void prefer_max_shared(K kernel) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

int launch_map_prepare(const float* feat, int64_t n, int32_t d, int32_t dpad, __nv_bfloat16* bf,
                       float* row_norm, float* row_c, float* row_an, float kappa, int f16, uint32_t* nonfinite,
                       int tiled, cudaStream_t s) {
  if (n == 0) return AVL_OK;
  const int warps = 8;
  const unsigned blocks = static_cast<unsigned>((n + warps - 1) / warps);
  map_prepare_kernel<<<blocks, warps * 32, 0, s>>>(feat, n, d, dpad, bf, row_norm, row_c, row_an, kappa, f16, nonfinite, tiled);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_query_prepare(const float* q, const float* fold_scale, int32_t nq, int32_t d, int32_t dpad,
                         int32_t npad, __nv_bfloat16* bq, float* q_bn, float* q_glob, int f16, cudaStream_t s) {
  // q_glob: [0] rho, [1] max ||b||, [2] ticket of the last-block reduction (zero at allocation, re-armed by the
  // kernel), [4 .. 4 + 256) per-row ratios
  prefer_max_shared(query_prepare_kernel);
  AVL_CUDA(launch_pdl(query_prepare_kernel, dim3(npad), dim3(128), 0, s, q, fold_scale, nq, d, dpad, bq, q_bn, q_glob + 4,
                      reinterpret_cast<uint32_t*>(q_glob), reinterpret_cast<uint32_t*>(q_glob) + 2, f16));
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_dense_exact(const float* feat, int64_t n, int32_t d, const float* q, int32_t nq,
                       const float* scale, const float* row_norm, int normalize, float* out,
                       int64_t out_rs, int64_t out_cs, cudaStream_t s) {
  if (n == 0 || nq == 0) return AVL_OK;
  dim3 grid(static_cast<unsigned>((n + kDT - 1) / kDT), static_cast<unsigned>((nq + kDT - 1) / kDT));
  dense_exact_kernel<<<grid, 256, 0, s>>>(feat, n, d, q, nq, scale, row_norm, normalize, out, out_rs, out_cs);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_column_exact(const float* feat, int64_t n, int32_t d, const float* q, const float* scale,
                        const float* row_norm, int normalize, float* out, int num_sms, cudaStream_t s) {
  if (n == 0) return AVL_OK;
  column_exact_kernel<<<num_sms * 8, 256, 0, s>>>(feat, n, d, q, scale, row_norm, normalize, out);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_argmax_rerank(const float* feat, int32_t d, const float* q, double* q64, const float* q_bn_raw, int32_t nq,
                         const float* scale,
                         const float* row_norm, int normalize, const uint32_t* flag_count,
                         const uint32_t* flag_rows, const uint32_t* flag_masks, uint32_t flag_cap,
                         int32_t* argmax_out, int num_sms, cudaStream_t s) {
  const int64_t nel = static_cast<int64_t>(nq) * d;
  f32_to_f64_kernel<<<static_cast<unsigned>((nel + 255) / 256), 256, 0, s>>>(q, q64, nel);
  argmax_rerank_kernel<<<num_sms * 8, 256, 0, s>>>(feat, d, q, q64, q_bn_raw, nq, scale, row_norm, normalize, flag_count,
                                                   flag_rows, flag_masks, flag_cap, argmax_out);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_select_threshold(const float* sample_lb, int32_t n_sample, int64_t ld, int32_t nq, int32_t k,
                            float* thr_t, cudaStream_t s) {
  if (k > kSelThreads) {
    set_error("select_threshold: k too large");
    return AVL_ERR_ARG;
  }
  prefer_max_shared(select_threshold_kernel);
  AVL_CUDA(launch_pdl(select_threshold_kernel, dim3(nq), dim3(kSelThreads), 0, s, sample_lb, n_sample, ld, k, thr_t));
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

size_t topk_finalize_smem(uint32_t cand_cap) { return static_cast<size_t>(cand_cap) * 12u; }

int launch_topk_finalize(const float* feat, int64_t n_rows, int32_t d, const float* q, int32_t nq,
                         const float* scale, const float* row_norm, const float* row_c, const float* row_an,
                         const float* q_bn, const float* q_glob, int normalize, int32_t k,
                         const uint32_t* bucket_cnt, int32_t grid, uint32_t cand_bucket, const uint32_t* cand_row,
                         const float* cand_val, uint32_t cand_cap, int64_t* out_idx, float* out_score,
                         uint32_t* cand_total, uint32_t* overflow_flags, uint32_t* gscratch, cudaStream_t s) {
  if (grid > kFinMaxGrid) {
    set_error("topk_finalize: screen grid larger than kFinMaxGrid");
    return AVL_ERR_UNSUPPORTED;
  }
  prefer_max_shared(topk_finalize_kernel);
  if (gscratch) {
    // beside the screen kernel: 128 threads (8 k registers per block: two blocks still fit next to a screen CTA's
    // 41.5 k -- with 256 threads the SMs that got two blocks had no room for the next call's sample-screen CTA, which
    // then waited for the finalize and the overlap was lost), no dynamic shared memory
    AVL_CUDA(launch_pdl(topk_finalize_kernel, dim3(nq), dim3(128), 0, s, feat, n_rows, d, q, scale, row_norm, row_c,
                        row_an, q_bn, reinterpret_cast<const uint32_t*>(q_glob), normalize, k, bucket_cnt, grid, cand_bucket,
                        cand_row, cand_val, cand_cap, out_idx, out_score, cand_total, overflow_flags, gscratch));
    return AVL_OK;
  }
  const size_t smem = topk_finalize_smem(cand_cap);
  AVL_CUDA(cudaFuncSetAttribute(topk_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem)));
  AVL_CUDA(launch_pdl(topk_finalize_kernel, dim3(nq), dim3(kFinThreads), smem, s, feat, n_rows, d, q, scale, row_norm, row_c,
                      row_an, q_bn, reinterpret_cast<const uint32_t*>(q_glob), normalize, k, bucket_cnt, grid, cand_bucket,
                      cand_row, cand_val, cand_cap, out_idx, out_score, cand_total, overflow_flags,
                      static_cast<uint32_t*>(nullptr)));
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

size_t topk_fallback_scratch_bytes(int num_sms) {
  return static_cast<size_t>(AVL_MAX_QUERIES) * (2 * num_sms) * AVL_MAX_TOPK * sizeof(unsigned long long);
}

int launch_topk_fallback(const float* feat, int64_t n, int32_t d, const float* q, int32_t nq, const float* scale,
                         const float* row_norm, int normalize, int32_t k, const uint32_t* overflow_flags,
                         void* scratch, uint32_t* tickets, int64_t* out_idx, float* out_score, int num_sms,
                         cudaStream_t s) {
  // queries per pass: bounded by 64 KiB of query vectors (fp64) and 40 KiB of per-warp lists in shared memory
  int group = kFbGroup;
  group = std::min<int>(group, std::max<int>(1, 65536 / (d * 8)));
  group = std::min<int>(group, std::max<int>(1, 40960 / ((kFbThreads / 32) * k * 8)));
  const size_t smem = static_cast<size_t>(group) * d * 8 + static_cast<size_t>(kFbThreads / 32) * group * k * 8;
  AVL_CUDA(cudaFuncSetAttribute(topk_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  prefer_max_shared(topk_fallback_kernel);
  AVL_CUDA(launch_pdl(topk_fallback_kernel, dim3(2 * num_sms), dim3(kFbThreads), smem, s, feat, n, d, q, nq, scale, row_norm,
                      normalize, k, group, overflow_flags, static_cast<unsigned long long*>(scratch), tickets, out_idx, out_score));
  return AVL_OK;
}

size_t topk_vector_scratch_bytes(int64_t n) {
  const int64_t blocks = (n + kVecChunk - 1) / kVecChunk;
  return static_cast<size_t>(blocks > 0 ? blocks : 1) * AVL_MAX_TOPK * sizeof(unsigned long long);
}

int launch_topk_vector(const float* values, int64_t n, int32_t k, int64_t* out_idx, float* out_val,
                       void* scratch, size_t scratch_bytes, cudaStream_t s) {
  const int64_t blocks = (n + kVecChunk - 1) / kVecChunk;
  if (n <= 0) {
    fill_u32_kernel<<<1, 256, 0, s>>>(reinterpret_cast<uint32_t*>(out_val), k, 0xFF800000u);
    AVL_CUDA(cudaMemsetAsync(out_idx, 0xFF, sizeof(int64_t) * k, s));
    return AVL_OK;
  }
  if (static_cast<size_t>(blocks) * k * sizeof(unsigned long long) > scratch_bytes) {
    set_error("topk_vector: scratch too small");
    return AVL_ERR_ARG;
  }
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(scratch);
  topk_vec_l1_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(values, n, k, keys);
  AVL_CUDA(cudaGetLastError());
  topk_vec_l2_kernel<<<1, 1024, 0, s>>>(keys, static_cast<int>(blocks * k), k, n, out_idx, out_val);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_merge_topk(const int64_t* idx, const float* val, int32_t n_shards, int32_t nq, int32_t k,
                      int64_t* out_idx, float* out_val, cudaStream_t s) {
  if (n_shards * k > 1024) {
    set_error("merge_topk: n_shards * k must be <= 1024");
    return AVL_ERR_UNSUPPORTED;
  }
  merge_topk_kernel<<<nq, 256, 0, s>>>(idx, val, n_shards, nq, k, out_idx, out_val);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_minmax_cols(const float* m, int64_t n, int32_t cols, float* out_min, float* out_max,
                       cudaStream_t s) {
  fill_u32_kernel<<<(cols + 255) / 256, 256, 0, s>>>(reinterpret_cast<uint32_t*>(out_min), cols, 0xFFFFFFFFu);
  fill_u32_kernel<<<(cols + 255) / 256, 256, 0, s>>>(reinterpret_cast<uint32_t*>(out_max), cols, 0u);
  dim3 grid(static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 592)), static_cast<unsigned>(cols));
  minmax_cols_kernel<<<grid, 256, 0, s>>>(m, n, reinterpret_cast<uint32_t*>(out_min),
                                          reinterpret_cast<uint32_t*>(out_max));
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

int launch_fuse_heat(const float* sa, const float* sb, int64_t n, int32_t pair, int32_t cols,
                     const float* min_a, const float* max_a, const float* min_b, const float* max_b,
                     int32_t combine, float* heat, cudaStream_t s) {
  (void)cols;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 1184));
  fuse_heat_kernel<<<blocks, 256, 0, s>>>(sa, sb, n, pair, reinterpret_cast<const uint32_t*>(min_a),
                                          reinterpret_cast<const uint32_t*>(max_a),
                                          reinterpret_cast<const uint32_t*>(min_b),
                                          reinterpret_cast<const uint32_t*>(max_b), combine, heat);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

// ---- fusion through the screen: the passes after the two dense screens (see the kernels above)
int launch_fuse_screened(const FuseSideHost& a, const FuseSideHost& b, int64_t n, int32_t pairs, int32_t combine,
                         int32_t k, const FuseScratch& w, int64_t* out_idx, float* out_heat, int num_sms,
                         cudaStream_t s) {
  auto side = [](const FuseSideHost& h) {
    FuseSide m;
    m.dense = h.dense; m.row_c = h.row_c; m.row_an = h.row_an; m.row_norm = h.row_norm; m.q_bn = h.q_bn;
    m.q_glob = h.q_glob; m.scale = h.scale; m.feat = h.feat; m.q = h.q; m.d = h.d; m.normalize = h.normalize;
    return m;
  };
  const FuseSide sa = side(a), sb = side(b);
  const int chunks = (pairs + kFuseChunk - 1) / kFuseChunk;
  // ~8 resident blocks per SM in total, at least 2 row-blocks per (chunk, side)
  const unsigned gx = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, std::max<int64_t>(2, static_cast<int64_t>(num_sms) * 8 / (2 * chunks))));
  AVL_CUDA(cudaMemsetAsync(w.max_lb, 0, sizeof(uint32_t) * 2 * pairs, s));
  AVL_CUDA(cudaMemsetAsync(w.min_ub, 0xFF, sizeof(uint32_t) * 2 * pairs, s));
  AVL_CUDA(cudaMemsetAsync(w.ext_cnt, 0, sizeof(uint32_t) * 4 * pairs, s));
  AVL_CUDA(cudaMemsetAsync(w.cand_cnt, 0, sizeof(uint32_t) * pairs, s));
  AVL_CUDA(cudaMemsetAsync(w.overflow, 0, sizeof(uint32_t), s));
  const bool vec = (n & 3) == 0;  // column bases stay 16-byte aligned
  if (vec) {
    fuse_colstats_kernel<4><<<dim3(gx, chunks, 2), 256, 0, s>>>(sa, sb, n, pairs, w.max_lb, w.min_ub);
    fuse_collect_extreme_kernel<4><<<dim3(gx, chunks, 2), 256, 0, s>>>(sa, sb, n, pairs, w.max_lb, w.min_ub, w.ext_cnt,
                                                                        w.ext_row, w.ext_cap);
  } else {
    fuse_colstats_kernel<1><<<dim3(gx, chunks, 2), 256, 0, s>>>(sa, sb, n, pairs, w.max_lb, w.min_ub);
    fuse_collect_extreme_kernel<1><<<dim3(gx, chunks, 2), 256, 0, s>>>(sa, sb, n, pairs, w.max_lb, w.min_ub, w.ext_cnt,
                                                                        w.ext_row, w.ext_cap);
  }
  fuse_exact_extreme_kernel<<<dim3(pairs, 2), 256, 0, s>>>(sa, sb, pairs, w.ext_cnt, w.ext_row, w.ext_cap, w.mm,
                                                           w.overflow);
  fuse_sample_kernel<<<(w.n_sample + 255) / 256, 256, 0, s>>>(sa, sb, n, pairs, w.sample_stride, w.n_sample, w.mm,
                                                              combine, w.sample_t);
  AVL_CUDA(cudaGetLastError());
  int rc = launch_select_threshold(w.sample_t, w.n_sample, w.n_sample, pairs, k, w.thr, s);
  if (rc) return rc;
  if (vec)
    fuse_collect_heat_kernel<4><<<dim3(2 * gx, chunks), 256, 0, s>>>(sa, sb, n, pairs, w.mm, combine, w.thr, w.cand_cnt,
                                                                     w.cand_row, w.cand_lo, w.cand_hi, w.cand_cap);
  else
    fuse_collect_heat_kernel<1><<<dim3(2 * gx, chunks), 256, 0, s>>>(sa, sb, n, pairs, w.mm, combine, w.thr, w.cand_cnt,
                                                                     w.cand_row, w.cand_lo, w.cand_hi, w.cand_cap);
  const size_t smem = static_cast<size_t>(w.cand_cap) * sizeof(uint32_t);
  AVL_CUDA(cudaFuncSetAttribute(fuse_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  fuse_finalize_kernel<<<pairs, 1024, smem, s>>>(sa, sb, pairs, w.mm, combine, k, w.cand_cnt, w.cand_row, w.cand_lo,
                                                 w.cand_hi, w.cand_cap, out_idx, out_heat, w.overflow);
  AVL_CUDA(cudaGetLastError());
  return AVL_OK;
}

}  // namespace avl
