// Map-build path: back-projection + alpha-weighted feature fusion, B200-native.
//
// Replaces the per-frame body of VLMapBuilder.create_mobile_base_map
// (reference avlmaps/map/vlmap_builder.py:129-178) and the helpers it calls per point
// (depth2pc mapping_utils.py:226-251, transform_pc :305-315, base_pos2grid_id_3d :345-349,
// project_point :599-605).  The reference loop is sequential and order dependent (first touch of a
// cell stores feat*alpha with weight alpha, later touches average); the order-free form used here
// (SURVEY.md section 0.4, appendix A):
//     key(point)   = (frame_seq << 32) | position in the frame's sample list
//     first(cell)  = min key over the points of the cell                     -> atomicMin
//     voxel id     = rank of first(cell) among all cells = running count of "winner" points in
//                    (frame, sample) order                                   -> ordered scan per frame
//     grid_feat    = (alpha_first^2 f_first + sum_{others} alpha f) / sum alpha
//     weight       = sum alpha
// Kernels per frame: geometry (fp64, the reference's exact operation sequence) -> winner count -> scan -> id assignment ->
// scatter-reduce (warp per point, float4 vector reds into the voxel row).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

#include <cuda_fp16.h>

#include "avl_internal.h"

namespace avl {
namespace {

struct FrameGeom {
  double kinv[9], k[9], kfeat[9], tf[16];
  double min_depth, max_depth, cs, half_gs;
  double origin[3];      // pcd_min (global-frame grid)
  int32_t h, w, fh, fw;
  int32_t n0, n1, n2;    // rows, cols, heights of occupied_ids
  int32_t mode;          // 0 = mobile-base grid (vlmap_builder.py), 1 = global-frame grid (vlmap_builder_multi_floor.py)
  int32_t slab_lo, slab_hi;  // rows owned by this builder
  int32_t depth_u16;     // depth is uint16 millimetres
  int32_t has_rgb;
};

// Up to kMaxBatch frames per launch triple (geometry, id scan, scatter): the samples of the frames are simply
// concatenated -- the first-touch key (frame_seq << 32 | position in the frame's sample list) already orders them.
// Passed by value as a __grid_constant__ kernel parameter (< 4 KiB), so a batch costs no upload.
constexpr int kMaxBatch = 16;  // 7.7 KiB of kernel parameters (CUDA >= 12.1 allows 32 KiB on sm_70+).  8 was enough on the GPUs
                               // next to the host cores; on the far socket of an 8-GPU box a launch costs ~50 us of
                               // host time, and 3 launches per 8 frames left those ranks of a slab build host-bound
struct FrameBatch {
  FrameGeom g[kMaxBatch];
  const float* depth[kMaxBatch];
  const int32_t* sidx[kMaxBatch];
  const float* feat[kMaxBatch];   // pixel-major (FH, FW, D)
  const uint8_t* rgb[kMaxBatch];
  int32_t off[kMaxBatch + 1];     // sample offsets of the frames inside the batch
  int32_t nf;
  uint32_t frame_seq0;            // frame_seq of g[0]
  int32_t feat_f16;               // feat[] point at __half rows (pixel-major fp16 hand-off), not float
};
static_assert(sizeof(FrameBatch) <= 16000, "FrameBatch must fit the kernel parameter space");

__device__ __forceinline__ int batch_frame_of(const FrameBatch& b, int gidx) {
  int f = 0;
#pragma unroll
  for (int i = 1; i < kMaxBatch; ++i) f += (i < b.nf && gidx >= b.off[i]) ? 1 : 0;
  return f;
}

constexpr int kScanBlock = 1024;
constexpr unsigned long long kNoKey = 0xFFFFFFFFFFFFFFFFull;

// The matrix products of the reference (cam_mat_inv @ p_2d, pose @ pc_homo, cam_mat @ p) run through
// numpy matmul -> OpenBLAS, whose x86-64 kernels accumulate over k ascending with fused multiply-adds:
// acc = m0*x; acc = fma(m1, y, acc); acc = fma(m2, z, acc).  This form reproduces numpy on every element
// (tools/probe_matmul_fma.py); the element-wise operations around them round separately.
__device__ __forceinline__ double dot3(const double* m, double x, double y, double z) {
  return __fma_rn(m[2], z, __fma_rn(m[1], y, __dmul_rn(m[0], x)));
}
__device__ __forceinline__ double dot4h(const double* m, double x, double y, double z) {  // row . [x, y, z, 1]
  return __fma_rn(m[3], 1.0, __fma_rn(m[2], z, __fma_rn(m[1], y, __dmul_rn(m[0], x))));
}
__device__ __forceinline__ long long trunc_ll(double v) {
  // python int(): toward zero; far-out values are clamped (the range tests reject them anyway)
  if (!(v > -9.0e15)) return -(1ll << 60);
  if (!(v < 9.0e15)) return (1ll << 60);
  return __double2ll_rz(v);
}
__device__ __forceinline__ long long round_ll(double v) {
  // np.round(...).astype(int): round half to even
  if (!(v > -9.0e15)) return -(1ll << 60);
  if (!(v < 9.0e15)) return (1ll << 60);
  return __double2ll_rn(v);
}
__device__ __forceinline__ double load_depth(const FrameGeom& g, const float* depth, int pix) {
  // multi-floor: load_depth_img(path) / 1000.0 -> uint16 / python float = float64 division
  // (vlmap_builder_multi_floor.py:103,128); mobile base: float32 .npy metres widened by numpy
  if (g.depth_u16) return __ddiv_rn(static_cast<double>(reinterpret_cast<const uint16_t*>(depth)[pix]), 1000.0);
  return static_cast<double>(depth[pix]);
}

// ---------------------------------------------------------------- geometry + first-touch keys
__global__ void __launch_bounds__(256)
geom_kernel(const __grid_constant__ FrameBatch batch, unsigned long long* __restrict__ first_key,
            int32_t* __restrict__ s_cell, int32_t* __restrict__ s_fpix, float* __restrict__ s_alpha,
            int32_t* __restrict__ s_rgbpix, uint8_t* __restrict__ s_wrap,
            unsigned long long* __restrict__ n_oob, uint32_t* __restrict__ ticket) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;  // for the look-back scan that follows in stream order
  const int n_total = batch.off[batch.nf];
  for (int gidx = blockIdx.x * blockDim.x + threadIdx.x; gidx < n_total; gidx += gridDim.x * blockDim.x) {
    const int fb = batch_frame_of(batch, gidx);
    const FrameGeom& g = batch.g[fb];
    const float* __restrict__ depth = batch.depth[fb];
    const int32_t* __restrict__ sample_idx = batch.sidx[fb];
    const uint32_t frame_seq = batch.frame_seq0 + static_cast<uint32_t>(fb);
    const int j = gidx - batch.off[fb];  // position in the frame's sample list
    const int pix = sample_idx ? sample_idx[j] : j;
    const int v = pix / g.w, u = pix - v * g.w;
    const double x2 = u + 0.5, y2 = v + 0.5;
    const double z = load_depth(g, depth, pix);
    // depth2pc: pc = (Kinv @ [u+.5, v+.5, 1]) * z   (mapping_utils.py:239-246)
    const double px = __dmul_rn(dot3(g.kinv + 0, x2, y2, 1.0), z);
    const double py = __dmul_rn(dot3(g.kinv + 3, x2, y2, 1.0), z);
    const double pz = __dmul_rn(dot3(g.kinv + 6, x2, y2, 1.0), z);
    int cell = -1, fpix = 0, rgbpix = -1;
    unsigned wrap = 0;
    float alpha = 0.f;
    if (pz > g.min_depth && pz < g.max_depth) {  // mapping_utils.py:247-249
      // transform_pc: pose @ [p; 1]   (mapping_utils.py:311-315)
      const double gx = dot4h(g.tf + 0, px, py, pz);
      const double gy = dot4h(g.tf + 4, px, py, pz);
      const double gz = dot4h(g.tf + 8, px, py, pz);
      long long row, col, hh;
      bool in_grid, height_oob = false;
      if (g.mode == 0) {
        // base_pos2grid_id_3d (mapping_utils.py:345-349): double truncation toward zero
        row = trunc_ll(__dsub_rn(g.half_gs, static_cast<double>(trunc_ll(__ddiv_rn(gx, g.cs)))));
        col = trunc_ll(__dsub_rn(g.half_gs, static_cast<double>(trunc_ll(__ddiv_rn(gy, g.cs)))));
        hh = trunc_ll(__ddiv_rn(gz, g.cs));
        in_grid = !(col >= g.n1 || row >= g.n0 || hh >= g.n2 || col < 0 || row < 0 || hh < 0);  // vlmap_builder.py:283
      } else {
        // row, height, col = np.round((p - pcd_min) / cs).astype(int)   (vlmap_builder_multi_floor.py:146)
        row = round_ll(__ddiv_rn(__dsub_rn(gx, g.origin[0]), g.cs));
        hh = round_ll(__ddiv_rn(__dsub_rn(gy, g.origin[1]), g.cs));
        col = round_ll(__ddiv_rn(__dsub_rn(gz, g.origin[2]), g.cs));
        in_grid = !(row >= g.n0 || col >= g.n1);  // the only test the reference makes (:151-153)
        if (in_grid) {
          // negative indices wrap like numpy's; grid_pos keeps the unwrapped values (:176)
          if (row < 0) { row += g.n0; wrap |= 1u; }
          if (col < 0) { col += g.n1; wrap |= 2u; }
          if (hh < 0) { hh += g.n2; wrap |= 4u; }
          if (row < 0 || col < 0) {  // height_map[row, col] raises IndexError in the reference (:155)
            in_grid = false;
            atomicAdd(n_oob, 1ull);
          }
          height_oob = hh < 0 || hh >= g.n2;  // occupied_ids[row, col, height] would raise (:175)
        }
      }
      if (in_grid && (row < g.slab_lo || row >= g.slab_hi)) in_grid = false;  // another rank's slab
      if (in_grid) {
        // project_point with the feature camera (vlmap_builder.py:143, mapping_utils.py:599-605)
        const double f2 = dot3(g.kfeat + 6, px, py, pz);
        const long long fx = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.kfeat + 0, px, py, pz), f2), 0.5));
        const long long fy = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.kfeat + 3, px, py, pz), f2), 0.5));
        if (!(fx < 0 || fy < 0 || fx >= g.fw || fy >= g.fh) && height_oob) atomicAdd(n_oob, 1ull);
        if (!(fx < 0 || fy < 0 || fx >= g.fw || fy >= g.fh) && !height_oob) {  // vlmap_builder.py:161
          cell = static_cast<int>((row * g.n1 + col) * g.n2 + hh);
          fpix = static_cast<int>(fy * g.fw + fx);
          // alpha = exp(-||p||^2 / (2 * 0.6))   (vlmap_builder.py:156-158)
          const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(px, px), __dmul_rn(py, py)), __dmul_rn(pz, pz));
          alpha = static_cast<float>(exp(__ddiv_rn(-rsq, 1.2)));
          if (g.has_rgb) {  // vlmap_builder.py:141-142: no bounds check; negative indices wrap like numpy
            const double q2 = dot3(g.k + 6, px, py, pz);
            long long rx = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.k + 0, px, py, pz), q2), 0.5));
            long long ry = trunc_ll(__dsub_rn(__ddiv_rn(dot3(g.k + 3, px, py, pz), q2), 0.5));
            if (rx < 0) rx += g.w;
            if (ry < 0) ry += g.h;
            if (rx >= 0 && rx < g.w && ry >= 0 && ry < g.h) rgbpix = static_cast<int>(ry * g.w + rx);
          }
          const unsigned long long key = (static_cast<unsigned long long>(frame_seq) << 32) | static_cast<uint32_t>(j);
          atomicMin(first_key + cell, key);
        }
      }
    }
    s_cell[gidx] = cell;
    s_fpix[gidx] = fpix;
    s_alpha[gidx] = alpha;
    s_rgbpix[gidx] = rgbpix;
    s_wrap[gidx] = static_cast<uint8_t>(wrap);
  }
}

// ---------------------------------------------------------------- pass 1 of the global-frame build
// min / max over the valid sampled points of transform_pc(depth2pc(depth), tf)
// (vlmap_builder_multi_floor.py:97-118); partial[block][0..2] = min xyz, [3..5] = max xyz, [6] = count
__global__ void __launch_bounds__(256)
bounds_kernel(const FrameGeom g, const float* __restrict__ depth, const int32_t* __restrict__ sample_idx,
              int32_t n_samples, double* __restrict__ partial) {
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  unsigned cnt = 0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n_samples; j += gridDim.x * blockDim.x) {
    const int pix = sample_idx ? sample_idx[j] : j;
    const int v = pix / g.w, u = pix - v * g.w;
    const double x2 = u + 0.5, y2 = v + 0.5;
    const double z = load_depth(g, depth, pix);
    const double px = __dmul_rn(dot3(g.kinv + 0, x2, y2, 1.0), z);
    const double py = __dmul_rn(dot3(g.kinv + 3, x2, y2, 1.0), z);
    const double pz = __dmul_rn(dot3(g.kinv + 6, x2, y2, 1.0), z);
    if (pz > g.min_depth && pz < g.max_depth) {
      const double gx = dot4h(g.tf + 0, px, py, pz);
      const double gy = dot4h(g.tf + 4, px, py, pz);
      const double gz = dot4h(g.tf + 8, px, py, pz);
      mn[0] = fmin(mn[0], gx); mn[1] = fmin(mn[1], gy); mn[2] = fmin(mn[2], gz);
      mx[0] = fmax(mx[0], gx); mx[1] = fmax(mx[1], gy); mx[2] = fmax(mx[2], gz);
      ++cnt;
    }
  }
  __shared__ double sh[8][7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) {
    for (int c = 0; c < 3; ++c) { sh[warp][c] = mn[c]; sh[warp][3 + c] = mx[c]; }
    sh[warp][6] = static_cast<double>(cnt);
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    const int c = threadIdx.x;
    double a = sh[0][c];
    for (int w2 = 1; w2 < 8; ++w2) a = c < 3 ? fmin(a, sh[w2][c]) : (c < 6 ? fmax(a, sh[w2][c]) : a + sh[w2][c]);
    partial[blockIdx.x * 7 + c] = a;
  }
}
// one block: acc[0..5] = min/max merged with the partials, acc[6] += count
__global__ void __launch_bounds__(32)
bounds_merge_kernel(const double* __restrict__ partial, int32_t nblocks, double* __restrict__ acc) {
  const int c = threadIdx.x;
  if (c >= 7) return;
  double a = acc[c];
  for (int b = 0; b < nblocks; ++b) {
    const double p = partial[b * 7 + c];
    a = c < 3 ? fmin(a, p) : (c < 6 ? fmax(a, p) : a + p);
  }
  acc[c] = a;
}

// keys[id] = first-touch key of voxel id; ids are assigned in key order, so keys come out ascending
__global__ void __launch_bounds__(256)
export_keys_kernel(const int32_t* __restrict__ occupied_ids, const unsigned long long* __restrict__ first_key,
                   int64_t cells, int64_t capacity, unsigned long long* __restrict__ keys) {
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < cells;
       c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t id = occupied_ids[c];
    if (id >= 0 && id < capacity) keys[id] = first_key[c];
  }
}

// resume: one warp per reloaded voxel -> num = grid_feat * weight, den = weight, cell -> id, key 0 (never a winner again)
__global__ void __launch_bounds__(256)
import_kernel(const float* __restrict__ feat, const int32_t* __restrict__ pos, const float* __restrict__ weight,
              const uint8_t* __restrict__ rgb, int64_t v, int32_t d, int32_t n0, int32_t n1, int32_t n2,
              float* __restrict__ num, float* __restrict__ den, float* __restrict__ rgb_acc,
              int32_t* __restrict__ grid_pos, int32_t* __restrict__ occupied_ids,
              unsigned long long* __restrict__ first_key, unsigned long long* __restrict__ n_bad) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t id = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; id < v; id += nwarps) {
    const float w = weight[id];
    for (int c = lane; c < d; c += 32) num[id * d + c] = feat[id * d + c] * w;
    if (lane < 3) {
      grid_pos[id * 3 + lane] = pos[id * 3 + lane];
      rgb_acc[id * 3 + lane] = rgb ? static_cast<float>(rgb[id * 3 + lane]) * w : 0.f;
    }
    if (lane == 0) {
      den[id] = w;
      long long r = pos[id * 3 + 0], c = pos[id * 3 + 1], hh = pos[id * 3 + 2];
      if (r < 0) r += n0;  // global-frame maps keep the unwrapped (negative) indices in grid_pos
      if (c < 0) c += n1;
      if (hh < 0) hh += n2;
      if (r < 0 || r >= n0 || c < 0 || c >= n1 || hh < 0 || hh >= n2) {
        atomicAdd(n_bad, 1ull);
      } else {
        const int64_t cell = (r * n1 + c) * n2 + hh;
        occupied_ids[cell] = static_cast<int32_t>(id);
        first_key[cell] = 0ull;
      }
    }
  }
}

constexpr int kMaxShards = 64;
struct ShardOffsets { int64_t off[kMaxShards + 1]; };
// global id = number of keys of all shards that are smaller (keys are unique: frame_seq << 32 | sample position)
__global__ void __launch_bounds__(256)
rank_keys_kernel(const unsigned long long* __restrict__ keys_all, const ShardOffsets so, int32_t n_shards,
                 int32_t shard, int64_t* __restrict__ out) {
  const int64_t n_own = so.off[shard + 1] - so.off[shard];
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < n_own;
       j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const unsigned long long key = keys_all[so.off[shard] + j];
    int64_t rank = j;
    for (int s2 = 0; s2 < n_shards; ++s2) {
      if (s2 == shard) continue;
      int64_t lo = so.off[s2], hi = so.off[s2 + 1];
      const int64_t base = lo;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (keys_all[mid] < key) lo = mid + 1; else hi = mid;
      }
      rank += lo - base;
    }
    out[j] = rank;
  }
}

__device__ __forceinline__ bool is_winner(const unsigned long long* first_key, int cell, uint32_t frame_seq, int j) {
  return cell >= 0 &&
         first_key[cell] == ((static_cast<unsigned long long>(frame_seq) << 32) | static_cast<uint32_t>(j));
}

__device__ __forceinline__ bool is_winner_b(const unsigned long long* first_key, int cell, const FrameBatch& batch, int gidx) {
  if (cell < 0) return false;
  const int fb = batch_frame_of(batch, gidx);
  return first_key[cell] == ((static_cast<unsigned long long>(batch.frame_seq0 + static_cast<uint32_t>(fb)) << 32) |
                             static_cast<uint32_t>(gidx - batch.off[fb]));
}

// ---------------------------------------------------------------- ordered id assignment
__global__ void __launch_bounds__(kScanBlock)
winner_count_kernel(const int32_t* __restrict__ s_cell, int32_t n_samples, uint32_t frame_seq,
                    const unsigned long long* __restrict__ first_key, uint32_t* __restrict__ block_cnt,
                    unsigned long long* __restrict__ n_accepted) {
  const int j = blockIdx.x * kScanBlock + threadIdx.x;
  const int cell = j < n_samples ? s_cell[j] : -1;
  const int wins = __syncthreads_count(is_winner(first_key, cell, frame_seq, j));
  const int acc = __syncthreads_count(cell >= 0);
  if (threadIdx.x == 0) {
    block_cnt[blockIdx.x] = static_cast<uint32_t>(wins);
    if (acc) atomicAdd(n_accepted, static_cast<unsigned long long>(acc));
  }
}

// one block: exclusive scan of block_cnt, offset by max_id; max_id += total
__global__ void __launch_bounds__(kScanBlock)
winner_scan_kernel(uint32_t* __restrict__ block_cnt, int32_t nblocks, unsigned long long* __restrict__ max_id) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = static_cast<uint32_t>(*max_id);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < nblocks; base += kScanBlock) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_cnt[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      uint32_t t = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += y;
      }
      warp_tot[lane] = t;  // inclusive totals of the warps
    }
    __syncthreads();
    const uint32_t before = carry + (warp ? warp_tot[warp - 1] : 0u) + (x - v);
    if (i < nblocks) block_cnt[i] = before;  // now: first id of the block
    __syncthreads();
    if (threadIdx.x == kScanBlock - 1) carry = before + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *max_id = carry;
}

__global__ void __launch_bounds__(kScanBlock)
assign_ids_kernel(const int32_t* __restrict__ s_cell, int32_t n_samples, uint32_t frame_seq,
                  const unsigned long long* __restrict__ first_key, const uint32_t* __restrict__ block_base,
                  int32_t n0, int32_t n1, int32_t n2, const uint8_t* __restrict__ s_wrap, int64_t capacity,
                  int32_t* __restrict__ occupied_ids, int32_t* __restrict__ grid_pos) {
  __shared__ uint32_t warp_tot[32];
  const int j = blockIdx.x * kScanBlock + threadIdx.x;
  const int cell = j < n_samples ? s_cell[j] : -1;
  const bool win = is_winner(first_key, cell, frame_seq, j);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = __ballot_sync(0xffffffffu, win);
  if (lane == 0) warp_tot[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    uint32_t t = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += y;
    }
    warp_tot[lane] = t;
  }
  __syncthreads();
  if (win) {
    const int64_t id = static_cast<int64_t>(block_base[blockIdx.x]) + (warp ? warp_tot[warp - 1] : 0u) +
                       __popc(b & ((1u << lane) - 1u));
    if (id < capacity) {
      occupied_ids[cell] = static_cast<int32_t>(id);  // vlmap_builder.py:165
      const int hh = cell % n2, rc = cell / n2;
      const unsigned wrap = s_wrap[j];                 // global-frame grid: indices that wrapped stay negative here
      grid_pos[id * 3 + 0] = rc / n1 - ((wrap & 1u) ? n0 : 0);  // vlmap_builder.py:169, vlmap_builder_multi_floor.py:176
      grid_pos[id * 3 + 1] = rc % n1 - ((wrap & 2u) ? n1 : 0);
      grid_pos[id * 3 + 2] = hh - ((wrap & 4u) ? n2 : 0);
    }
  }
}

// ---------------------------------------------------------------- ordered id assignment, single pass
// winner count + exclusive scan + id assignment in ONE kernel (decoupled look-back): block t (ticket order)
// publishes its winner count, then sums the counts of its predecessors until it meets an inclusive prefix.
// scan_state[t] = frame tag (30 bits) | status (2 bits: 1 = aggregate, 2 = inclusive prefix) | value (32 bits);
// the tag makes entries of earlier frames read as "not yet published", so the array is never cleared.
// *ticket is zeroed by the geometry kernel of the same frame (stream order).
constexpr unsigned long long kStAggregate = 1ull << 32, kStInclusive = 2ull << 32;

__global__ void __launch_bounds__(kScanBlock)
assign_ids_lookback_kernel(const __grid_constant__ FrameBatch batch, const int32_t* __restrict__ s_cell,
                           const unsigned long long* __restrict__ first_key, unsigned long long* __restrict__ scan_state,
                           uint32_t* __restrict__ ticket, int32_t n0, int32_t n1, int32_t n2,
                           const uint8_t* __restrict__ s_wrap, int64_t capacity, int32_t* __restrict__ occupied_ids,
                           int32_t* __restrict__ grid_pos, unsigned long long* __restrict__ max_id,
                           unsigned long long* __restrict__ n_accepted) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t s_ticket, s_base;
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t t = s_ticket;
  const int n_samples = batch.off[batch.nf];
  const uint32_t frame_seq = batch.frame_seq0;  // tag of this launch's scan state
  const int j = static_cast<int>(t) * kScanBlock + threadIdx.x;
  const int cell = j < n_samples ? s_cell[j] : -1;
  const bool win = is_winner_b(first_key, cell, batch, j);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t b = __ballot_sync(0xffffffffu, win);
  const uint32_t acc_b = __ballot_sync(0xffffffffu, cell >= 0);
  if (lane == 0) warp_tot[warp] = static_cast<uint32_t>(__popc(b)) | (static_cast<uint32_t>(__popc(acc_b)) << 16);
  __syncthreads();
  if (warp == 0) {
    const uint32_t packed = warp_tot[lane];
    uint32_t wins = packed & 0xffffu, acc = packed >> 16;
    uint32_t incl = wins;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    warp_tot[lane] = incl;  // inclusive winner totals of the warps
    const uint32_t agg = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned long long tag = static_cast<unsigned long long>(frame_seq & 0x3fffffffu) << 34;
    uint32_t prefix = 0;
    if (t == 0) {
      prefix = static_cast<uint32_t>(*max_id);
    } else {
      if (lane == 0) {
        *reinterpret_cast<volatile unsigned long long*>(scan_state + t) = tag | kStAggregate | agg;
        __threadfence();
      }
      int idx = static_cast<int>(t) - 1;
      while (true) {
        const int k = idx - lane;
        unsigned long long v = 0;
        if (k >= 0) {
          do {
            v = *reinterpret_cast<volatile unsigned long long*>(scan_state + k);
          } while ((v >> 34) != (tag >> 34) || ((v >> 32) & 3ull) == 0ull);
        } else {
          v = tag | kStInclusive;  // before block 0: contributes nothing, terminates the walk
        }
        const bool inclusive = ((v >> 32) & 3ull) == 2ull;
        const uint32_t first = __ffs(__ballot_sync(0xffffffffu, inclusive));  // nearest predecessor with a full prefix
        uint32_t val = (first == 0 || lane < static_cast<int>(first)) ? static_cast<uint32_t>(v) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
        prefix += val;
        if (first) break;
        idx -= 32;
      }
    }
    if (lane == 0) {
      *reinterpret_cast<volatile unsigned long long*>(scan_state + t) = tag | kStInclusive | (prefix + agg);
      __threadfence();
      s_base = prefix;
      if (acc) atomicAdd(n_accepted, static_cast<unsigned long long>(acc));
      if (t == gridDim.x - 1) *max_id = prefix + agg;
    }
  }
  __syncthreads();
  if (win) {
    const int64_t id = static_cast<int64_t>(s_base) + (warp ? warp_tot[warp - 1] : 0u) + __popc(b & ((1u << lane) - 1u));
    if (id < capacity) {
      occupied_ids[cell] = static_cast<int32_t>(id);  // vlmap_builder.py:165
      const int hh = cell % n2, rc = cell / n2;
      const unsigned wrap = s_wrap[j];
      grid_pos[id * 3 + 0] = rc / n1 - ((wrap & 1u) ? n0 : 0);  // vlmap_builder.py:169, vlmap_builder_multi_floor.py:176
      grid_pos[id * 3 + 1] = rc % n1 - ((wrap & 2u) ? n1 : 0);
      grid_pos[id * 3 + 2] = hh - ((wrap & 4u) ? n2 : 0);
    }
  }
}

// ---------------------------------------------------------------- feature layout
// (D, P) -> (P, D), P = FH*FW pixels.  64 pixels x 64 channels per block through shared memory:
// 256-byte contiguous reads per channel row, 256-byte contiguous writes per pixel row (float4 both
// ways when P and D allow it).
__global__ void __launch_bounds__(256)
chw_to_hwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int32_t d, int64_t p) {
  __shared__ float tile[64][65];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  const bool vec_in = (p & 3) == 0, vec_out = (d & 3) == 0;
  // load: row = channel, 16 threads x float4 cover 64 pixels
  {
    const int q4 = (t & 15) * 4, r0 = t >> 4;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 16) {
      const int c = c0 + r0 + rr;
      const int64_t pp = p0 + q4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        const float* s = src + static_cast<int64_t>(c) * p + pp;
        if (vec_in && pp + 3 < p) v = __ldg(reinterpret_cast<const float4*>(s));
        else {
          if (pp < p) v.x = s[0];
          if (pp + 1 < p) v.y = s[1];
          if (pp + 2 < p) v.z = s[2];
          if (pp + 3 < p) v.w = s[3];
        }
      }
      tile[r0 + rr][q4] = v.x; tile[r0 + rr][q4 + 1] = v.y; tile[r0 + rr][q4 + 2] = v.z; tile[r0 + rr][q4 + 3] = v.w;
    }
  }
  __syncthreads();
  // store: row = pixel, 16 threads x float4 cover 64 channels
  {
    const int c4 = (t & 15) * 4, r0 = t >> 4;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 16) {
      const int64_t pp = p0 + r0 + rr;
      const int c = c0 + c4;
      if (pp >= p) continue;
      const float4 v = make_float4(tile[c4][r0 + rr], tile[c4 + 1][r0 + rr], tile[c4 + 2][r0 + rr], tile[c4 + 3][r0 + rr]);
      float* o = dst + pp * d + c;
      if (vec_out && c + 3 < d) *reinterpret_cast<float4*>(o) = v;
      else {
        if (c < d) o[0] = v.x;
        if (c + 1 < d) o[1] = v.y;
        if (c + 2 < d) o[2] = v.z;
        if (c + 3 < d) o[3] = v.w;
      }
    }
  }
}

// Same transposition for fp16 features (AVL_FEAT_F16): (D, P) __half -> (P, D) float.  LSeg emits
// `logit_scale * normalize(feat).half()` (lseg_net.py:318-321), so the fp32 array get_lseg_feat hands over holds
// fp16-exact values: taking the halves directly halves the hand-off bytes (PCIe or HBM) and changes no result.
__global__ void __launch_bounds__(256)
chw16_to_hwc_kernel(const __half* __restrict__ src, float* __restrict__ dst, int32_t d, int64_t p) {
  __shared__ float tile[64][65];
  const int64_t p0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  const bool vec_in = (p & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 7) == 0, vec_out = (d & 3) == 0;
  {
    const int q4 = (t & 15) * 4, r0 = t >> 4;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 16) {
      const int c = c0 + r0 + rr;
      const int64_t pp = p0 + q4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < d) {
        const __half* s = src + static_cast<int64_t>(c) * p + pp;
        if (vec_in && pp + 3 < p) {  // 4 halves = one 8-byte load (c * p + pp is a multiple of 4)
          const uint2 raw = __ldg(reinterpret_cast<const uint2*>(s));
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          v = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          if (pp < p) v.x = __half2float(s[0]);
          if (pp + 1 < p) v.y = __half2float(s[1]);
          if (pp + 2 < p) v.z = __half2float(s[2]);
          if (pp + 3 < p) v.w = __half2float(s[3]);
        }
      }
      tile[r0 + rr][q4] = v.x; tile[r0 + rr][q4 + 1] = v.y; tile[r0 + rr][q4 + 2] = v.z; tile[r0 + rr][q4 + 3] = v.w;
    }
  }
  __syncthreads();
  {
    const int c4 = (t & 15) * 4, r0 = t >> 4;
#pragma unroll
    for (int rr = 0; rr < 64; rr += 16) {
      const int64_t pp = p0 + r0 + rr;
      const int c = c0 + c4;
      if (pp >= p) continue;
      const float4 v = make_float4(tile[c4][r0 + rr], tile[c4 + 1][r0 + rr], tile[c4 + 2][r0 + rr], tile[c4 + 3][r0 + rr]);
      float* o = dst + pp * d + c;
      if (vec_out && c + 3 < d) *reinterpret_cast<float4*>(o) = v;
      else {
        if (c < d) o[0] = v.x;
        if (c + 1 < d) o[1] = v.y;
        if (c + 2 < d) o[2] = v.z;
        if (c + 3 < d) o[3] = v.w;
      }
    }
  }
}

// ---------------------------------------------------------------- scatter-reduce
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// One warp per accepted point: num[id, :] += w * feat[fpix, :], den[id] += alpha, rgb likewise.
// w = alpha^2 for the point that first touched the cell (vlmap_builder.py:164-170 stores feat*alpha
// with weight alpha), alpha otherwise (:171-178).
template <bool kChunk>
__global__ void __launch_bounds__(256)
scatter_kernel(const __grid_constant__ FrameBatch batch, int32_t d,
               const int32_t* __restrict__ s_cell, const int32_t* __restrict__ s_fpix,
               const float* __restrict__ s_alpha, const int32_t* __restrict__ s_rgbpix,
               const unsigned long long* __restrict__ first_key,
               const int32_t* __restrict__ occupied_ids, int64_t capacity, float* __restrict__ num,
               float* __restrict__ den, float* __restrict__ rgb_acc) {
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_samples = batch.off[batch.nf];
  // A warp takes 32 consecutive samples at a time: one coalesced load of their cells and voxel ids, a ballot of the
  // accepted ones, then one pass of the whole warp per accepted sample.  (One sample per warp iteration made every
  // REJECTED sample cost a dependent load: two thirds of a frame's samples on one GPU, and 15 of 16 in an 8-way
  // slab-sharded build, where that walk -- not the scatter -- set the frame time.)
  // kChunk = false walks one sample per warp step (every lane holds that sample): the finer interleave is ~8 % faster
  // when a launch is a single frame on one GPU (about one 32-sample chunk per warp: no averaging over chunks).
  constexpr int kStep = kChunk ? 32 : 1;
  for (int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kStep; base < n_samples; base += nwarps * kStep) {
    const int jl = kChunk ? base + lane : base;
    const int cell_l = jl < n_samples ? s_cell[jl] : -1;
    const int64_t id_l = cell_l >= 0 ? static_cast<int64_t>(occupied_ids[cell_l]) : -1;
    const bool ok_l = id_l >= 0 && id_l < capacity;
    // everything else a sample needs is loaded by its own lane too (coalesced, in flight together) and handed to the
    // warp by shuffles: no dependent broadcast load is left inside the per-sample pass
    float alpha_l = 0.f;
    int fpix_l = 0, rgbpix_l = -1, win_l = 0;
    if (ok_l) {
      alpha_l = s_alpha[jl];
      fpix_l = s_fpix[jl];
      rgbpix_l = s_rgbpix[jl];
      const int fbl = batch_frame_of(batch, jl);
      win_l = first_key[cell_l] == ((static_cast<unsigned long long>(batch.frame_seq0 + static_cast<uint32_t>(fbl)) << 32) |
                                    static_cast<uint32_t>(jl - batch.off[fbl]));
    }
    uint32_t todo = kChunk ? __ballot_sync(0xffffffffu, ok_l) : (ok_l ? 1u : 0u);
    while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const int j = kChunk ? base + src : base;
    const int64_t id = __shfl_sync(0xffffffffu, id_l, src);
    const float alpha = __shfl_sync(0xffffffffu, alpha_l, src);
    const int fpix = __shfl_sync(0xffffffffu, fpix_l, src);
    const int fb = batch_frame_of(batch, j);
    const bool win = __shfl_sync(0xffffffffu, win_l, src) != 0;
    const float wgt = win ? alpha * alpha : alpha;
    const float* __restrict__ feat_hwc = batch.feat[fb];
    const uint8_t* __restrict__ rgb = batch.rgb[fb];
    float* o = num + id * d;
    if (batch.feat_f16) {
      // pixel-major fp16 rows, what LSeg itself emits (lseg_net.py:318-321: `.half()`): the feature read of a point
      // is D * 2 bytes instead of D * 4 -- 5 144 instead of 6 168 bytes per accepted point -- and the values are the
      // very same (fp16 -> fp32 is exact)
      const __half* fh = reinterpret_cast<const __half*>(feat_hwc) + static_cast<int64_t>(fpix) * d;
      if ((d & 3) == 0) {
        const uint2* f2 = reinterpret_cast<const uint2*>(fh);
        for (int c = lane; c < (d >> 2); c += 32) {
          const uint2 raw = __ldg(f2 + c);
          const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
          const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
          red_add_v4(o + 4 * c, lo.x * wgt, lo.y * wgt, hi.x * wgt, hi.y * wgt);
        }
      } else {
        for (int c = lane; c < d; c += 32) atomicAdd(o + c, __half2float(fh[c]) * wgt);
      }
    } else {
    const float* f = feat_hwc + static_cast<int64_t>(fpix) * d;
    if ((d & 3) == 0) {
      const float4* f4 = reinterpret_cast<const float4*>(f);
      for (int c = lane; c < (d >> 2); c += 32) {
        const float4 v = __ldg(f4 + c);
        red_add_v4(o + 4 * c, v.x * wgt, v.y * wgt, v.z * wgt, v.w * wgt);
      }
    } else {
      for (int c = lane; c < d; c += 32) atomicAdd(o + c, f[c] * wgt);
    }
    }
    if (lane == 0) atomicAdd(den + id, alpha);
    const int rp_s = __shfl_sync(0xffffffffu, rgbpix_l, src);
    if (rgb && lane >= 1 && lane <= 3) {
      const int rp = rp_s;
      const float cv = rp >= 0 ? static_cast<float>(rgb[static_cast<int64_t>(rp) * 3 + (lane - 1)]) : 0.f;
      atomicAdd(rgb_acc + id * 3 + (lane - 1), cv * alpha);
    }
    }
  }
}

// ---------------------------------------------------------------- export
__global__ void __launch_bounds__(256)
export_feat_kernel(const float* __restrict__ num, const float* __restrict__ den, int64_t v, int32_t d,
                   float* __restrict__ out) {
  const int64_t total = v * d;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = __fdiv_rn(num[i], den[i / d]);
}
__global__ void __launch_bounds__(256)
export_rgb_kernel(const float* __restrict__ rgb_acc, const float* __restrict__ den, int64_t v,
                  uint8_t* __restrict__ out) {
  const int64_t total = v * 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float x = __fdiv_rn(rgb_acc[i], den[i / 3]);
    out[i] = static_cast<uint8_t>(fminf(fmaxf(x, 0.f), 255.f));  // truncation, like the uint8 store
  }
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}
__global__ void fill_u64_kernel(unsigned long long* p, int64_t n, unsigned long long v) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = v;
}

}  // namespace
}  // namespace avl

using namespace avl;

struct avl_builder {
  int32_t n0 = 0, n1 = 0, n2 = 0;  // occupied_ids dims: rows, cols, heights
  int32_t dim = 0;
  double cs = 0.0;
  int32_t mode = 0;                // 0 mobile-base grid, 1 global-frame (multi-floor) grid
  double origin[3] = {0.0, 0.0, 0.0};
  int32_t slab_lo = 0, slab_hi = 0;
  int num_sms = 148;
  int64_t cells = 0;
  int64_t capacity = 0;
  uint32_t frame_seq = 0;
  int64_t id_upper = 0;  // host-side upper bound of max_id (avoids a sync per frame)
  // persistent device state
  unsigned long long* first_key = nullptr;
  int32_t* occupied_ids = nullptr;
  float* num = nullptr;
  float* den = nullptr;
  float* rgb_acc = nullptr;
  int32_t* grid_pos = nullptr;
  unsigned long long* counters = nullptr;  // [0] max_id, [1] n_accepted, [2] rejected where the reference raises
  // per-frame scratch (grown on demand)
  int32_t *s_cell = nullptr, *s_fpix = nullptr, *s_rgbpix = nullptr;
  float* s_alpha = nullptr;
  uint8_t* s_wrap = nullptr;
  uint32_t* block_cnt = nullptr;
  unsigned long long* scan_state = nullptr;  // per 1024-sample block: look-back scan state (tagged by frame)
  uint32_t* ticket = nullptr;
  int64_t scratch_samples = 0;
  // staging of host inputs
  float* d_depth = nullptr; size_t depth_elems = 0;
  float* d_feat = nullptr; size_t feat_elems = 0;
  float* d_feat_t = nullptr; size_t feat_t_elems = 0;
  uint8_t* d_rgb = nullptr; size_t rgb_bytes = 0;
  int32_t* d_sidx = nullptr; size_t sidx_elems = 0;
  // sparse hand-off of host-resident (1, D, FH, FW) features (add_frame_sparse): pinned host buffers
  int32_t* h_cell = nullptr; int32_t* h_fpix = nullptr; size_t h_samples = 0;   // geometry results read back
  void* h_stage[2] = {nullptr, nullptr}; size_t h_stage_bytes[2] = {0, 0};      // gathered channel rows, double-buffered
  int32_t* h_cidx[2] = {nullptr, nullptr}; size_t h_cidx_elems[2] = {0, 0};     // compact feature index per sample
  cudaEvent_t h_ev[2] = {nullptr, nullptr}; bool h_busy[2] = {false, false};
  int h_next = 0;
  std::vector<int32_t> h_rank, h_upix;
  uint64_t h2d_bytes = 0;  // bytes uploaded from host pointers so far (avl_builder_h2d_bytes)
};

namespace {

// ---------------------------------------------------------------- host threads (feature gather)
// A small persistent pool: the gather of a frame at depth_sample_rate 100 is ~0.1 ms of work, so threads are not
// spawned per call.  run(n, f) calls f(i) for i in [0, n) on the pool and the caller.
class HostPool {
 public:
  static HostPool& get() {
    static HostPool p;
    return p;
  }
  int size() const { return static_cast<int>(th_.size()) + 1; }
  void run(int n, const std::function<void(int)>& f) {
    if (n <= 0) return;
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &f;
      n_ = n;
      next_.store(0);
      left_ = n;
      ++gen_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return left_ == 0; });
    job_ = nullptr;
  }

 private:
  HostPool() {
    // the CPUs this process may run on (a container's cpuset), not the machine's: hardware_concurrency() reports the
    // latter and oversubscribed a 16-CPU cpuset four times over
    int hc = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) hc = CPU_COUNT(&set);
    if (hc <= 0) hc = static_cast<int>(std::thread::hardware_concurrency());
    const char* e = getenv("AVL_HOST_THREADS");
    int n = e ? atoi(e) : (hc > 0 ? hc : 4);
    n = std::max(1, std::min(n, 64));
    for (int i = 1; i < n; ++i) th_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void work() {
    for (;;) {
      const int i = next_.fetch_add(1);
      if (i >= n_) break;
      (*job_)(i);
      std::lock_guard<std::mutex> lk(m_);
      if (--left_ == 0) done_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      work();
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* job_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, left_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// Upload from PAGEABLE host memory (a numpy array) at PCIe speed: cudaMemcpyAsync from pageable memory goes through the
// driver's single-threaded bounce buffer (~10 GB/s measured on B200's host); here the host threads copy 16 MiB chunks
// into two pinned buffers and each chunk is handed to the copy engine as soon as it is staged.  Returns when the
// source has been read.
struct StagedUpload {
  uint8_t* buf[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  bool busy[2] = {false, false};
  static constexpr size_t kChunk = size_t(16) << 20;
  int copy(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s) {
    for (int i = 0; i < 2; ++i) {
      if (!buf[i]) AVL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&buf[i]), kChunk, cudaHostAllocDefault));
      if (!ev[i]) AVL_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
    const int nt = HostPool::get().size();
    int slot = 0;
    for (size_t off = 0; off < bytes; off += kChunk, slot ^= 1) {
      const size_t n = std::min(kChunk, bytes - off);
      if (busy[slot]) {
        AVL_CUDA(cudaEventSynchronize(ev[slot]));
        busy[slot] = false;
      }
      const uint8_t* src = static_cast<const uint8_t*>(src_host) + off;
      uint8_t* dst = buf[slot];
      const size_t piece = (n + nt - 1) / nt;
      HostPool::get().run(nt, [&](int t) {
        const size_t b0 = static_cast<size_t>(t) * piece;
        if (b0 < n) memcpy(dst + b0, src + b0, std::min(piece, n - b0));
      });
      AVL_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(dst_dev) + off, dst, n, cudaMemcpyHostToDevice, s));
      AVL_CUDA(cudaEventRecord(ev[slot], s));
      busy[slot] = true;
    }
    return AVL_OK;
  }
  ~StagedUpload() {}  // process teardown: see HeatScratch
};
static thread_local StagedUpload g_staged_upload;

template <typename T>
int grow_pinned(T** p, size_t* have, size_t need) {
  if (*have >= need) return AVL_OK;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *have = 0;
  AVL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(p), need * sizeof(T), cudaHostAllocDefault));
  *have = need;
  return AVL_OK;
}

template <typename T>
int grow(T** p, size_t* have, size_t need) {
  if (*have >= need) return AVL_OK;
  cudaFree(*p);
  *p = nullptr;
  *have = 0;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(p), need * sizeof(T)));
  *have = need;
  return AVL_OK;
}

int alloc_rows(avl_builder* b, int64_t cap, float** num, float** den, float** rgb, int32_t** pos, cudaStream_t s) {
  const int d = b->dim;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(num), static_cast<size_t>(cap) * d * sizeof(float)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(den), static_cast<size_t>(cap) * sizeof(float)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(rgb), static_cast<size_t>(cap) * 3 * sizeof(float)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(pos), static_cast<size_t>(cap) * 3 * sizeof(int32_t)));
  AVL_CUDA(cudaMemsetAsync(*num, 0, static_cast<size_t>(cap) * d * sizeof(float), s));
  AVL_CUDA(cudaMemsetAsync(*den, 0, static_cast<size_t>(cap) * sizeof(float), s));
  AVL_CUDA(cudaMemsetAsync(*rgb, 0, static_cast<size_t>(cap) * 3 * sizeof(float), s));
  AVL_CUDA(cudaMemsetAsync(*pos, 0, static_cast<size_t>(cap) * 3 * sizeof(int32_t), s));
  return AVL_OK;
}

// capacity doubling, the analogue of _reserve_map_space (vlmap_builder.py:286-311)
int ensure_capacity(avl_builder* b, int64_t incoming, cudaStream_t s) {
  // a builder can only ever create voxels in the cells of its own row slab (all rows unless avl_builder_set_slab)
  const int64_t own_cells = static_cast<int64_t>(b->slab_hi - b->slab_lo) * b->n1 * b->n2;
  if (b->capacity >= own_cells) return AVL_OK;  // one row per own cell already: cannot overflow
  if (b->id_upper + incoming <= b->capacity) {
    b->id_upper += incoming;
    return AVL_OK;
  }
  unsigned long long max_id = 0;
  AVL_CUDA(cudaMemcpyAsync(&max_id, b->counters, sizeof(max_id), cudaMemcpyDeviceToHost, s));
  AVL_CUDA(cudaStreamSynchronize(s));
  b->id_upper = static_cast<int64_t>(max_id);
  if (b->id_upper + incoming > b->capacity) {
    int64_t cap = b->capacity;
    while (cap < b->id_upper + incoming) cap *= 2;
    if (cap > own_cells) cap = own_cells;  // never more rows than cells that can be occupied
    float *num, *den, *rgb;
    int32_t* pos;
    int rc = alloc_rows(b, cap, &num, &den, &rgb, &pos, s);
    if (rc) return rc;
    const size_t v = static_cast<size_t>(b->id_upper), d = b->dim;
    AVL_CUDA(cudaMemcpyAsync(num, b->num, v * d * sizeof(float), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaMemcpyAsync(den, b->den, v * sizeof(float), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaMemcpyAsync(rgb, b->rgb_acc, v * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaMemcpyAsync(pos, b->grid_pos, v * 3 * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    AVL_CUDA(cudaStreamSynchronize(s));
    cudaFree(b->num); cudaFree(b->den); cudaFree(b->rgb_acc); cudaFree(b->grid_pos);
    b->num = num; b->den = den; b->rgb_acc = rgb; b->grid_pos = pos;
    b->capacity = cap;
  }
  b->id_upper += incoming;
  return AVL_OK;
}

void fill_geom(FrameGeom* g, const avl_frame* f, int flags) {
  memset(g, 0, sizeof(*g));
  memcpy(g->kinv, f->kinv, sizeof(g->kinv));
  memcpy(g->k, f->k, sizeof(g->k));
  memcpy(g->kfeat, f->kfeat, sizeof(g->kfeat));
  memcpy(g->tf, f->tf, sizeof(g->tf));
  g->min_depth = f->min_depth;
  g->max_depth = f->max_depth;
  g->h = f->h; g->w = f->w; g->fh = f->fh; g->fw = f->fw;
  g->depth_u16 = (flags & AVL_DEPTH_U16_MM) ? 1 : 0;
}


struct BatchItem {
  const avl_frame* f;
  const float* depth; const float* feat; const uint8_t* rgb; const int32_t* sidx;  // device pointers, feat pixel-major
  int32_t n_samples;
  bool feat_f16 = false;   // feat points at pixel-major __half rows
};

// per-sample scratch of one launch triple (cell, feature pixel, colour pixel, alpha, wrap flag) + the scan's block state;
// grown (never shrunk) to the largest batch seen, or ahead of time through avl_builder_reserve
int grow_scratch(avl_builder* b, int64_t n_samples, cudaStream_t s) {
  if (b->scratch_samples >= n_samples) return AVL_OK;
  cudaFree(b->s_cell); cudaFree(b->s_fpix); cudaFree(b->s_rgbpix); cudaFree(b->s_alpha); cudaFree(b->s_wrap);
  cudaFree(b->block_cnt); cudaFree(b->scan_state);
  b->s_cell = b->s_fpix = b->s_rgbpix = nullptr; b->s_alpha = nullptr; b->s_wrap = nullptr; b->block_cnt = nullptr;
  b->scan_state = nullptr;
  b->scratch_samples = 0;
  const size_t n = static_cast<size_t>(n_samples);
  const size_t nb = (n + kScanBlock - 1) / kScanBlock;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_cell), n * sizeof(int32_t)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_fpix), n * sizeof(int32_t)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_rgbpix), n * sizeof(int32_t)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_alpha), n * sizeof(float)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->s_wrap), n));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->block_cnt), nb * sizeof(uint32_t)));
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->scan_state), nb * sizeof(unsigned long long)));
  AVL_CUDA(cudaMemsetAsync(b->scan_state, 0xff, nb * sizeof(unsigned long long), s));
  if (!b->ticket) AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->ticket), sizeof(uint32_t)));
  b->scratch_samples = n_samples;
  return AVL_OK;
}

// geometry -> ordered id scan -> scatter for up to kMaxBatch frames whose inputs are on the device
// phase: 1 = geometry only (first-touch keys, per-sample cell / feature pixel / alpha), 2 = id scan + scatter of a
// batch whose geometry ran, 3 = both.  The split lets a host-resident frame learn WHICH feature pixels it needs
// (phase 1), gather only those on the host, and fuse (phase 2).
int launch_batch(avl_builder* b, const BatchItem* items, int nf, int flags, cudaStream_t s, int phase = 3) {
  FrameBatch batch;
  memset(&batch, 0, sizeof(batch));
  int64_t total = 0;
  for (int i = 0; i < nf; ++i) {
    batch.off[i] = static_cast<int32_t>(total);
    total += items[i].n_samples;
  }
  if (total >= (int64_t(1) << 31)) {
    set_error("batch has more than 2^31 samples");
    return AVL_ERR_ARG;
  }
  batch.off[nf] = static_cast<int32_t>(total);
  batch.nf = nf;
  batch.frame_seq0 = b->frame_seq;
  batch.feat_f16 = (nf > 0 && items[0].feat_f16) ? 1 : 0;
  if (total == 0) {
    if (phase & 2) b->frame_seq += static_cast<uint32_t>(nf);
    return AVL_OK;
  }
  const int32_t n_samples = static_cast<int32_t>(total);
  int rc;
  // ---- scratch
  if ((phase & 1) && (rc = grow_scratch(b, n_samples, s))) return rc;
  if ((phase & 1) && (rc = ensure_capacity(b, n_samples, s))) return rc;

  for (int i = 0; i < nf; ++i) {
    FrameGeom& g = batch.g[i];
    fill_geom(&g, items[i].f, flags);
    g.cs = b->cs;
    g.half_gs = b->n0 / 2.0;
    memcpy(g.origin, b->origin, sizeof(g.origin));
    g.n0 = b->n0; g.n1 = b->n1; g.n2 = b->n2;
    g.mode = b->mode;
    g.slab_lo = b->slab_lo; g.slab_hi = b->slab_hi;
    g.has_rgb = items[i].rgb != nullptr;
    batch.depth[i] = items[i].depth;
    batch.sidx[i] = items[i].sidx;
    batch.feat[i] = items[i].feat;
    batch.rgb[i] = items[i].rgb;
  }
  const int d = b->dim;
  const int nblocks = (n_samples + kScanBlock - 1) / kScanBlock;
  const int geom_blocks = std::min((n_samples + 255) / 256, b->num_sms * 8);
  if (phase & 1)
    geom_kernel<<<geom_blocks, 256, 0, s>>>(batch, b->first_key, b->s_cell, b->s_fpix, b->s_alpha, b->s_rgbpix, b->s_wrap,
                                            b->counters + 2, b->ticket);
  if (!(phase & 2)) {
    AVL_CUDA(cudaGetLastError());
    return AVL_OK;
  }
  static const bool three_pass = getenv("AVL_BUILD_3PASS") != nullptr;  // the original count / scan / assign kernels (A/B)
  if (three_pass && nf == 1) {
    winner_count_kernel<<<nblocks, kScanBlock, 0, s>>>(b->s_cell, n_samples, b->frame_seq, b->first_key,
                                                       b->block_cnt, b->counters + 1);
    winner_scan_kernel<<<1, kScanBlock, 0, s>>>(b->block_cnt, nblocks, b->counters);
    assign_ids_kernel<<<nblocks, kScanBlock, 0, s>>>(b->s_cell, n_samples, b->frame_seq, b->first_key, b->block_cnt,
                                                     b->n0, b->n1, b->n2, b->s_wrap, b->capacity, b->occupied_ids,
                                                     b->grid_pos);
  } else {
    assign_ids_lookback_kernel<<<nblocks, kScanBlock, 0, s>>>(batch, b->s_cell, b->first_key, b->scan_state, b->ticket,
                                                              b->n0, b->n1, b->n2, b->s_wrap, b->capacity,
                                                              b->occupied_ids, b->grid_pos, b->counters, b->counters + 1);
  }
  const int scatter_blocks = std::min((n_samples + 7) / 8, b->num_sms * 8);
  // 32-sample chunks per warp step when a warp gets several of them, or when this builder owns a row slab only (most
  // samples are then another rank's: the chunked walk skips 32 of them per step)
  const bool chunked = (b->slab_hi - b->slab_lo) < b->n0 || static_cast<int64_t>(n_samples) > int64_t(2) * scatter_blocks * 8 * 32;
  if (chunked)
    scatter_kernel<true><<<scatter_blocks, 256, 0, s>>>(batch, d, b->s_cell, b->s_fpix, b->s_alpha, b->s_rgbpix, b->first_key,
                                                        b->occupied_ids, b->capacity, b->num, b->den, b->rgb_acc);
  else
    scatter_kernel<false><<<scatter_blocks, 256, 0, s>>>(batch, d, b->s_cell, b->s_fpix, b->s_alpha, b->s_rgbpix, b->first_key,
                                                         b->occupied_ids, b->capacity, b->num, b->den, b->rgb_acc);
  AVL_CUDA(cudaGetLastError());
  b->frame_seq += static_cast<uint32_t>(nf);
  return AVL_OK;
}

// Host-resident (1, D, FH, FW) features, the array get_lseg_feat returns (lseg_utils.py:101-102): 415 MB of float32
// per 390 x 520 x 512 frame, of which the frame's accepted points read at most 100 k pixel rows (205 MB) at
// depth_sample_rate 1 and ~3 000 (6 MB) at the reference's default rate 100 (config/map_config/vlmaps.yaml:13).
// Copying the whole array made the call PCIe-bound at 127 frames/s whatever the rate.  Here the geometry runs first and
// tells the host WHICH feature pixels are read; the host gathers exactly those, channel row by channel row (one
// contiguous, sorted pass over each 811 KB channel plane per thread), into pinned staging; the GPU transposes the
// compact (D, n_used) block and fuses.  The call returns once the caller's arrays have been consumed: the fusion of
// frame i overlaps the host gather of frame i + 1 (staging is double-buffered).
template <typename T>
static void gather_channel_rows(const T* src, size_t plane, const int32_t* upix, int32_t nu, T* dst, int d) {
  HostPool::get().run(d, [&](int c) {
    const T* p = src + static_cast<size_t>(c) * plane;
    T* o = dst + static_cast<size_t>(c) * nu;
    for (int32_t u = 0; u < nu; ++u) o[u] = p[upix[u]];
  });
}

int add_frame_sparse(avl_builder* b, const avl_frame* f, const float* depth_dev, const uint8_t* rgb_dev,
                     const int32_t* sidx_dev, int32_t n_samples, int flags, cudaStream_t s) {
  const int d = b->dim;
  const size_t fpix = static_cast<size_t>(f->fh) * f->fw;
  const bool f16 = (flags & AVL_FEAT_F16) != 0;
  const size_t esz = f16 ? sizeof(__half) : sizeof(float);
  int rc;
  BatchItem it;
  it.f = f;
  it.depth = depth_dev; it.feat = nullptr; it.rgb = rgb_dev; it.sidx = sidx_dev;
  it.n_samples = n_samples;
  // ---- 1. geometry on the device, cells and feature pixels back
  if ((rc = launch_batch(b, &it, 1, flags, s, /*phase=*/1))) return rc;
  if (b->h_samples < static_cast<size_t>(n_samples)) {
    if (b->h_cell) cudaFreeHost(b->h_cell);
    if (b->h_fpix) cudaFreeHost(b->h_fpix);
    b->h_cell = b->h_fpix = nullptr;
    b->h_samples = 0;
    AVL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b->h_cell), n_samples * sizeof(int32_t), cudaHostAllocDefault));
    AVL_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&b->h_fpix), n_samples * sizeof(int32_t), cudaHostAllocDefault));
    b->h_samples = n_samples;
  }
  AVL_CUDA(cudaMemcpyAsync(b->h_cell, b->s_cell, n_samples * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  AVL_CUDA(cudaMemcpyAsync(b->h_fpix, b->s_fpix, n_samples * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  AVL_CUDA(cudaStreamSynchronize(s));
  // ---- 2. the feature pixels in use, ascending, and every sample's row in the compact block
  const int slot = b->h_next;
  b->h_next ^= 1;
  if (b->h_busy[slot]) {  // the frame that last used this staging slot (two calls ago) must have been uploaded
    AVL_CUDA(cudaEventSynchronize(b->h_ev[slot]));
    b->h_busy[slot] = false;
  }
  if (!b->h_ev[slot]) AVL_CUDA(cudaEventCreateWithFlags(&b->h_ev[slot], cudaEventDisableTiming));
  b->h_rank.assign(fpix, 0);
  int32_t* rank = b->h_rank.data();
  for (int32_t j = 0; j < n_samples; ++j)
    if (b->h_cell[j] >= 0) rank[b->h_fpix[j]] = 1;   // accepted points have a feature pixel inside the map (vlmap_builder.py:161)
  b->h_upix.clear();
  for (size_t p = 0; p < fpix; ++p) {
    if (rank[p]) {
      rank[p] = static_cast<int32_t>(b->h_upix.size());
      b->h_upix.push_back(static_cast<int32_t>(p));
    }
  }
  const int32_t nu = static_cast<int32_t>(b->h_upix.size());
  // A frame at depth_sample_rate 1 reads ~40 % of the feature pixels: gathering them touches nearly every cache line
  // of the 415 MB array on the host and costs what copying all of it over PCIe costs (measured on B200: 10.7 vs 7.9 ms
  // per frame), so beyond a quarter of the pixels the whole array is uploaded; at the reference's default rate 100 the
  // frame reads 1.5 % of the pixels and the gather wins 12x (0.63 ms).
  if (static_cast<size_t>(nu) * 4 > fpix) {
    if ((rc = grow(&b->d_feat, &b->feat_elems, fpix * d))) return rc;
    if ((rc = grow(&b->d_feat_t, &b->feat_t_elems, fpix * d))) return rc;
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, f->feat) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();   // an unregistered pointer is not an error here
    if (pinned) AVL_CUDA(cudaMemcpyAsync(b->d_feat, f->feat, fpix * d * esz, cudaMemcpyHostToDevice, s));   // PCIe speed as it is
    else if ((rc = g_staged_upload.copy(b->d_feat, f->feat, fpix * d * esz, s))) return rc;
    b->h2d_bytes += fpix * d * esz;
    dim3 grid(static_cast<unsigned>((fpix + 63) / 64), static_cast<unsigned>((d + 63) / 64));
    if (f16)
      chw16_to_hwc_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(b->d_feat), b->d_feat_t, d, static_cast<int64_t>(fpix));
    else
      chw_to_hwc_kernel<<<grid, 256, 0, s>>>(b->d_feat, b->d_feat_t, d, static_cast<int64_t>(fpix));
    AVL_CUDA(cudaGetLastError());
    it.feat = b->d_feat_t;   // s_fpix on the device still holds the original feature pixels
    return launch_batch(b, &it, 1, flags, s, /*phase=*/2);
  }
  // ---- 3. gather the channel rows of those pixels into pinned staging
  if ((rc = grow_pinned(&b->h_cidx[slot], &b->h_cidx_elems[slot], static_cast<size_t>(n_samples)))) return rc;
  int32_t* cidx = b->h_cidx[slot];
  for (int32_t j = 0; j < n_samples; ++j) cidx[j] = b->h_cell[j] >= 0 ? rank[b->h_fpix[j]] : 0;
  const size_t need = std::max<size_t>(static_cast<size_t>(nu) * d * esz, 16);
  if (b->h_stage_bytes[slot] < need) {
    if (b->h_stage[slot]) cudaFreeHost(b->h_stage[slot]);
    b->h_stage[slot] = nullptr;
    b->h_stage_bytes[slot] = 0;
    const size_t cap = std::max(need, std::min(fpix / 4 + 1, static_cast<size_t>(n_samples)) * d * esz);  // its largest possible size
    AVL_CUDA(cudaHostAlloc(&b->h_stage[slot], cap, cudaHostAllocDefault));
    b->h_stage_bytes[slot] = cap;
  }
  if (nu > 0) {
    if (f16) gather_channel_rows(reinterpret_cast<const uint16_t*>(f->feat), fpix, b->h_upix.data(), nu,
                                 static_cast<uint16_t*>(b->h_stage[slot]), d);
    else gather_channel_rows(f->feat, fpix, b->h_upix.data(), nu, static_cast<float*>(b->h_stage[slot]), d);
  }
  // ---- 4. upload the compact block, transpose it to pixel-major rows, fuse
  const size_t nu1 = static_cast<size_t>(std::max(nu, 1));
  if ((rc = grow(&b->d_feat, &b->feat_elems, nu1 * d))) return rc;
  if ((rc = grow(&b->d_feat_t, &b->feat_t_elems, nu1 * d))) return rc;
  AVL_CUDA(cudaMemcpyAsync(b->s_fpix, cidx, n_samples * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  b->h2d_bytes += static_cast<uint64_t>(n_samples) * sizeof(int32_t) + static_cast<uint64_t>(nu) * d * esz;
  if (nu > 0) {
    AVL_CUDA(cudaMemcpyAsync(b->d_feat, b->h_stage[slot], static_cast<size_t>(nu) * d * esz, cudaMemcpyHostToDevice, s));
    dim3 grid(static_cast<unsigned>((nu + 63) / 64), static_cast<unsigned>((d + 63) / 64));
    if (f16)
      chw16_to_hwc_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(b->d_feat), b->d_feat_t, d, static_cast<int64_t>(nu));
    else
      chw_to_hwc_kernel<<<grid, 256, 0, s>>>(b->d_feat, b->d_feat_t, d, static_cast<int64_t>(nu));
    AVL_CUDA(cudaGetLastError());
  }
  AVL_CUDA(cudaEventRecord(b->h_ev[slot], s));
  b->h_busy[slot] = true;
  it.feat = b->d_feat_t;
  return launch_batch(b, &it, 1, flags, s, /*phase=*/2);
}

}  // namespace

struct avl_bounds {
  int num_sms = 148;
  double* acc = nullptr;      // [7] min xyz, max xyz, count
  double* partial = nullptr;  // [blocks][7]
  float* d_depth = nullptr; size_t depth_elems = 0;
  int32_t* d_sidx = nullptr; size_t sidx_elems = 0;
};

extern "C" {

static int builder_create_common(int32_t n0, int32_t n1, int32_t n2, double cs, int32_t dim, int64_t capacity,
                                 int32_t mode, const double* origin, avl_builder** out) {
  const int64_t cells = static_cast<int64_t>(n0) * n1 * n2;
  AVL_ARG(cells < (int64_t(1) << 31), "grid has more than 2^31 cells");
  int dev = 0, major = 0;
  AVL_CUDA(cudaGetDevice(&dev));
  AVL_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) {
    set_error("avlmaps_b200 needs an sm_100a (B200) device");
    return AVL_ERR_UNSUPPORTED;
  }
  avl_builder* b = new avl_builder();
  b->n0 = n0; b->n1 = n1; b->n2 = n2; b->cs = cs; b->dim = dim; b->mode = mode;
  if (origin) memcpy(b->origin, origin, sizeof(b->origin));
  b->slab_lo = 0;
  b->slab_hi = n0;
  cudaDeviceGetAttribute(&b->num_sms, cudaDevAttrMultiProcessorCount, dev);
  b->cells = cells;
  b->capacity = capacity > 0 ? std::min<int64_t>(capacity, cells)
                             : std::min<int64_t>(static_cast<int64_t>(n0) * n1, cells);
  int rc = AVL_OK;
  do {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&b->first_key), cells * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->occupied_ids), cells * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->counters), 4 * sizeof(unsigned long long));
    if (e != cudaSuccess) { rc = cuda_fail(e, "builder state", __FILE__, __LINE__); break; }
    fill_u64_kernel<<<1184, 256>>>(b->first_key, cells, kNoKey);
    fill_i32_kernel<<<1184, 256>>>(b->occupied_ids, cells, -1);  // vlmap_builder.py:204
    e = cudaMemset(b->counters, 0, 4 * sizeof(unsigned long long));
    if (e != cudaSuccess) { rc = cuda_fail(e, "builder init", __FILE__, __LINE__); break; }
    if ((rc = alloc_rows(b, b->capacity, &b->num, &b->den, &b->rgb_acc, &b->grid_pos, nullptr))) break;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { rc = cuda_fail(e, "builder init", __FILE__, __LINE__); break; }
  } while (0);
  if (rc) {
    avl_builder_destroy(b);
    return rc;
  }
  *out = b;
  return AVL_OK;
}

int avl_builder_create(const avl_grid_spec* spec, avl_builder** out) {
  AVL_ARG(spec != nullptr && out != nullptr, "NULL argument");
  *out = nullptr;
  AVL_ARG(spec->gs >= 1 && spec->vh >= 1 && spec->dim >= 1 && spec->cs > 0.0, "invalid grid spec");
  return builder_create_common(spec->gs, spec->gs, spec->vh, spec->cs, spec->dim, spec->capacity, 0, nullptr, out);
}

int avl_builder_create_global(const avl_global_grid_spec* spec, avl_builder** out) {
  AVL_ARG(spec != nullptr && out != nullptr, "NULL argument");
  *out = nullptr;
  AVL_ARG(spec->n_row >= 1 && spec->n_col >= 1 && spec->n_height >= 1 && spec->dim >= 1 && spec->cs > 0.0,
          "invalid grid spec");
  return builder_create_common(spec->n_row, spec->n_col, spec->n_height, spec->cs, spec->dim, spec->capacity, 1,
                               spec->pcd_min, out);
}

int avl_builder_set_slab(avl_builder* b, int32_t row_lo, int32_t row_hi) {
  AVL_ARG(b != nullptr, "builder is NULL");
  AVL_ARG(row_lo >= 0 && row_lo <= row_hi && row_hi <= b->n0, "slab outside the grid");
  if (b->frame_seq != 0) {
    set_error("avl_builder_set_slab must be called before the first frame");
    return AVL_ERR_STATE;
  }
  b->slab_lo = row_lo;
  b->slab_hi = row_hi;
  return AVL_OK;
}

int avl_builder_skip_frames(avl_builder* b, int32_t n_frames) {
  AVL_ARG(b != nullptr && n_frames >= 0, "invalid argument");
  b->frame_seq += static_cast<uint32_t>(n_frames);  // the skipped frames keep their place in the (frame, sample) order
  return AVL_OK;
}

int avl_builder_reserve(avl_builder* b, int64_t samples_per_call, void* stream) {
  AVL_ARG(b != nullptr && samples_per_call >= 0 && samples_per_call < (int64_t(1) << 31), "invalid argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int rc = grow_scratch(b, samples_per_call, s);
  if (rc) return rc;
  AVL_CUDA(cudaStreamSynchronize(s));  // a set-up call: the scan state is initialised before ANY stream's first frame
  return AVL_OK;
}

int avl_builder_destroy(avl_builder* b) {
  if (!b) return AVL_OK;
  cudaFree(b->first_key); cudaFree(b->occupied_ids); cudaFree(b->num); cudaFree(b->den); cudaFree(b->rgb_acc);
  cudaFree(b->grid_pos); cudaFree(b->counters); cudaFree(b->s_cell); cudaFree(b->s_fpix); cudaFree(b->s_rgbpix);
  cudaFree(b->s_alpha); cudaFree(b->s_wrap); cudaFree(b->block_cnt); cudaFree(b->scan_state); cudaFree(b->ticket);
  cudaFree(b->d_depth); cudaFree(b->d_feat); cudaFree(b->d_feat_t);
  cudaFree(b->d_rgb); cudaFree(b->d_sidx);
  if (b->h_cell) cudaFreeHost(b->h_cell);
  if (b->h_fpix) cudaFreeHost(b->h_fpix);
  for (int i = 0; i < 2; ++i) {
    if (b->h_stage[i]) cudaFreeHost(b->h_stage[i]);
    if (b->h_cidx[i]) cudaFreeHost(b->h_cidx[i]);
    if (b->h_ev[i]) cudaEventDestroy(b->h_ev[i]);
  }
  delete b;
  return AVL_OK;
}

int avl_builder_add_frame(avl_builder* b, const avl_frame* f, int flags, void* stream) {
  AVL_ARG(b != nullptr && f != nullptr, "NULL argument");
  AVL_ARG(f->depth != nullptr && f->feat != nullptr, "depth / feat is NULL");
  AVL_ARG(f->h >= 1 && f->w >= 1 && f->fh >= 1 && f->fw >= 1, "invalid frame shape");
  AVL_ARG(static_cast<int64_t>(f->h) * f->w < (int64_t(1) << 31), "frame too large");
  AVL_ARG(f->feat_layout == AVL_FEAT_CHW || f->feat_layout == AVL_FEAT_HWC, "unknown feat_layout");
  const bool feat_f16 = (flags & AVL_FEAT_F16) != 0;
  const int64_t npix = static_cast<int64_t>(f->h) * f->w;
  const int32_t n_samples = f->sample_idx ? f->n_samples : static_cast<int32_t>(npix);
  AVL_ARG(n_samples >= 0 && n_samples <= npix, "n_samples out of range");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int d = b->dim;
  const size_t fpix = static_cast<size_t>(f->fh) * f->fw;
  const size_t depth_bytes = static_cast<size_t>(npix) * ((flags & AVL_DEPTH_U16_MM) ? sizeof(uint16_t) : sizeof(float));
  int rc;

  // ---- inputs on the device
  const float* depth = f->depth;
  const float* feat = f->feat;
  const uint8_t* rgb = f->rgb;
  const int32_t* sidx = f->sample_idx;
  // host-resident channel-major features: only the feature pixels the frame's accepted points read are handed over
  static const bool dense_h2d = getenv("AVL_BUILD_DENSE_H2D") != nullptr;  // A/B: copy the whole (1, D, FH, FW) array
  const bool sparse = !(flags & AVL_ON_DEVICE) && f->feat_layout == AVL_FEAT_CHW && !dense_h2d && n_samples > 0;
  if (!(flags & AVL_ON_DEVICE)) {
    if ((rc = grow(&b->d_depth, &b->depth_elems, static_cast<size_t>(npix)))) return rc;
    AVL_CUDA(cudaMemcpyAsync(b->d_depth, f->depth, depth_bytes, cudaMemcpyHostToDevice, s));
    depth = b->d_depth;
    b->h2d_bytes += depth_bytes + (f->rgb ? static_cast<uint64_t>(npix) * 3 : 0) +
                    (f->sample_idx ? static_cast<uint64_t>(n_samples) * sizeof(int32_t) : 0);
    if (!sparse) {
      b->h2d_bytes += fpix * d * (feat_f16 ? sizeof(__half) : sizeof(float));  // d_feat is sized in floats: halves fit
      if ((rc = grow(&b->d_feat, &b->feat_elems, fpix * d))) return rc;
      AVL_CUDA(cudaMemcpyAsync(b->d_feat, f->feat, fpix * d * (feat_f16 ? sizeof(__half) : sizeof(float)),
                               cudaMemcpyHostToDevice, s));
      feat = b->d_feat;
    }
    if (f->rgb) {
      if ((rc = grow(&b->d_rgb, &b->rgb_bytes, static_cast<size_t>(npix) * 3))) return rc;
      AVL_CUDA(cudaMemcpyAsync(b->d_rgb, f->rgb, npix * 3, cudaMemcpyHostToDevice, s));
      rgb = b->d_rgb;
    }
    if (f->sample_idx) {
      if ((rc = grow(&b->d_sidx, &b->sidx_elems, static_cast<size_t>(std::max(n_samples, 1))))) return rc;
      AVL_CUDA(cudaMemcpyAsync(b->d_sidx, f->sample_idx, static_cast<size_t>(n_samples) * sizeof(int32_t),
                               cudaMemcpyHostToDevice, s));
      sidx = b->d_sidx;
    }
  }
  if (n_samples == 0) {
    b->frame_seq++;
    return AVL_OK;
  }
  if (sparse) return add_frame_sparse(b, f, depth, rgb, sidx, n_samples, flags, s);
  if (f->feat_layout == AVL_FEAT_CHW) {  // (1, D, FH, FW) -> pixel-major rows for the coalesced gather
    if ((rc = grow(&b->d_feat_t, &b->feat_t_elems, fpix * d))) return rc;
    dim3 grid(static_cast<unsigned>((fpix + 63) / 64), static_cast<unsigned>((d + 63) / 64));
    if (feat_f16)
      chw16_to_hwc_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(feat), b->d_feat_t, d, static_cast<int64_t>(fpix));
    else
      chw_to_hwc_kernel<<<grid, 256, 0, s>>>(feat, b->d_feat_t, d, static_cast<int64_t>(fpix));
    AVL_CUDA(cudaGetLastError());
    feat = b->d_feat_t;
  }

  // one frame = a batch of one (device pointers, pixel-major features by now)
  BatchItem it;
  it.f = f;
  it.depth = depth; it.feat = feat; it.rgb = rgb; it.sidx = sidx;
  it.n_samples = n_samples;
  it.feat_f16 = feat_f16 && f->feat_layout == AVL_FEAT_HWC;   // channel-major halves were widened by the transposition
  if ((rc = launch_batch(b, &it, 1, flags, s))) return rc;
  if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaStreamSynchronize(s));  // staging buffers are reused per frame
  return AVL_OK;
}

int avl_builder_add_frames(avl_builder* b, const avl_frame* frames, int32_t n_frames, int flags, void* stream) {
  AVL_ARG(b != nullptr && (frames != nullptr || n_frames == 0), "NULL argument");
  AVL_ARG(n_frames >= 0, "n_frames < 0");
  bool batchable = (flags & AVL_ON_DEVICE) != 0;
  for (int i = 0; i < n_frames && batchable; ++i) batchable = frames[i].feat_layout == AVL_FEAT_HWC;
  if (!batchable) {  // host pointers (staged per frame) or channel-major features (transposed per frame): one by one
    for (int i = 0; i < n_frames; ++i) {
      const int rc = avl_builder_add_frame(b, frames + i, flags, stream);
      if (rc) return rc;
    }
    return AVL_OK;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  for (int i0 = 0; i0 < n_frames; i0 += kMaxBatch) {
    BatchItem items[kMaxBatch];
    const int nb = std::min(kMaxBatch, n_frames - i0);
    int used = 0;
    for (int i = 0; i < nb; ++i) {
      const avl_frame* f = frames + i0 + i;
      AVL_ARG(f->depth != nullptr && f->feat != nullptr, "depth / feat is NULL");
      AVL_ARG(f->h >= 1 && f->w >= 1 && f->fh >= 1 && f->fw >= 1, "invalid frame shape");
      const int64_t npix = static_cast<int64_t>(f->h) * f->w;
      AVL_ARG(npix < (int64_t(1) << 28), "frame too large for a batch");
      const int32_t n_samples = f->sample_idx ? f->n_samples : static_cast<int32_t>(npix);
      AVL_ARG(n_samples >= 0 && n_samples <= npix, "n_samples out of range");
      BatchItem& it = items[used++];
      it.f = f;
      it.depth = f->depth; it.feat = f->feat; it.rgb = f->rgb; it.sidx = f->sample_idx;
      it.n_samples = n_samples;  // frames without samples stay in the batch: they still consume a frame_seq
      it.feat_f16 = (flags & AVL_FEAT_F16) != 0;   // pixel-major fp16 rows, read as they are by the scatter
    }
    const int rc = launch_batch(b, items, used, flags, s);
    if (rc) return rc;
  }
  return AVL_OK;
}

static int read_counter(avl_builder* b, int which, int64_t* n, void* stream) {
  AVL_ARG(b != nullptr && n != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long v = 0;
  AVL_CUDA(cudaMemcpyAsync(&v, b->counters + which, sizeof(v), cudaMemcpyDeviceToHost, s));
  AVL_CUDA(cudaStreamSynchronize(s));
  *n = static_cast<int64_t>(v);
  return AVL_OK;
}
int avl_builder_num_voxels(avl_builder* b, int64_t* n, void* stream) { return read_counter(b, 0, n, stream); }
int avl_builder_num_accepted(avl_builder* b, int64_t* n, void* stream) { return read_counter(b, 1, n, stream); }
int avl_builder_h2d_bytes(avl_builder* b, int64_t* n) {
  AVL_ARG(b != nullptr && n != nullptr, "NULL argument");
  *n = static_cast<int64_t>(b->h2d_bytes);
  return AVL_OK;
}
int avl_builder_num_rejected_oob(avl_builder* b, int64_t* n, void* stream) { return read_counter(b, 2, n, stream); }

int avl_builder_export(avl_builder* b, float* grid_feat, int32_t* grid_pos, float* weight, int32_t* occupied_ids,
                       uint8_t* grid_rgb, int flags, void* stream) {
  AVL_ARG(b != nullptr, "builder is NULL");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t v = 0;
  int rc = read_counter(b, 0, &v, stream);
  if (rc) return rc;
  const int d = b->dim;
  const cudaMemcpyKind kind = (flags & AVL_ON_DEVICE) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  if (grid_feat && v > 0) {
    float* dst = grid_feat;
    if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dst), static_cast<size_t>(v) * d * sizeof(float)));
    export_feat_kernel<<<b->num_sms * 8, 256, 0, s>>>(b->num, b->den, v, d, dst);
    if (!(flags & AVL_ON_DEVICE)) {
      cudaError_t e = cudaMemcpyAsync(grid_feat, dst, static_cast<size_t>(v) * d * sizeof(float), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      cudaFree(dst);
      if (e != cudaSuccess) return cuda_fail(e, "export grid_feat", __FILE__, __LINE__);
    }
  }
  if (grid_rgb && v > 0) {
    uint8_t* dst = grid_rgb;
    if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dst), static_cast<size_t>(v) * 3));
    export_rgb_kernel<<<b->num_sms * 2, 256, 0, s>>>(b->rgb_acc, b->den, v, dst);
    if (!(flags & AVL_ON_DEVICE)) {
      cudaError_t e = cudaMemcpyAsync(grid_rgb, dst, static_cast<size_t>(v) * 3, cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      cudaFree(dst);
      if (e != cudaSuccess) return cuda_fail(e, "export grid_rgb", __FILE__, __LINE__);
    }
  }
  if (grid_pos && v > 0) AVL_CUDA(cudaMemcpyAsync(grid_pos, b->grid_pos, static_cast<size_t>(v) * 3 * sizeof(int32_t), kind, s));
  if (weight && v > 0) AVL_CUDA(cudaMemcpyAsync(weight, b->den, static_cast<size_t>(v) * sizeof(float), kind, s));
  if (occupied_ids) AVL_CUDA(cudaMemcpyAsync(occupied_ids, b->occupied_ids, static_cast<size_t>(b->cells) * sizeof(int32_t), kind, s));
  AVL_CUDA(cudaGetLastError());
  if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaStreamSynchronize(s));
  return AVL_OK;
}

int avl_builder_import(avl_builder* b, const float* grid_feat, const int32_t* grid_pos, const float* weight,
                       const uint8_t* grid_rgb, int64_t n_voxels, int flags, void* stream) {
  AVL_ARG(b != nullptr, "builder is NULL");
  AVL_ARG(n_voxels >= 0 && n_voxels <= b->cells, "n_voxels exceeds the number of cells");
  if (b->frame_seq != 0 || b->id_upper != 0) {
    set_error("avl_builder_import needs a fresh builder (no frame added yet)");
    return AVL_ERR_STATE;
  }
  b->frame_seq = 1;  // keys of the frames to come are > 0 = the key of every reloaded cell
  if (n_voxels == 0) return AVL_OK;
  AVL_ARG(grid_feat != nullptr && grid_pos != nullptr && weight != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t v = static_cast<size_t>(n_voxels), d = b->dim;
  if (b->capacity < n_voxels) {  // rows for the reloaded voxels (the reference's arrays are exactly this long, :221)
    cudaFree(b->num); cudaFree(b->den); cudaFree(b->rgb_acc); cudaFree(b->grid_pos);
    b->num = b->den = b->rgb_acc = nullptr; b->grid_pos = nullptr;
    int rc = alloc_rows(b, n_voxels, &b->num, &b->den, &b->rgb_acc, &b->grid_pos, s);
    if (rc) return rc;
    b->capacity = n_voxels;
  }
  const float* f = grid_feat; const int32_t* p = grid_pos; const float* w = weight; const uint8_t* c = grid_rgb;
  float *df = nullptr, *dw = nullptr; int32_t* dp = nullptr; uint8_t* dc = nullptr;
  cudaError_t e = cudaSuccess;
  if (!(flags & AVL_ON_DEVICE)) {
    e = cudaMalloc(reinterpret_cast<void**>(&df), v * d * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&dp), v * 3 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&dw), v * sizeof(float));
    if (e == cudaSuccess && grid_rgb) e = cudaMalloc(reinterpret_cast<void**>(&dc), v * 3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(df, grid_feat, v * d * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dp, grid_pos, v * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dw, weight, v * sizeof(float), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && grid_rgb) e = cudaMemcpyAsync(dc, grid_rgb, v * 3, cudaMemcpyHostToDevice, s);
    f = df; p = dp; w = dw; c = grid_rgb ? dc : nullptr;
  }
  unsigned long long n_bad = 0;
  if (e == cudaSuccess) e = cudaMemsetAsync(b->counters + 3, 0, sizeof(unsigned long long), s);
  if (e == cudaSuccess) {
    import_kernel<<<b->num_sms * 8, 256, 0, s>>>(f, p, w, c, n_voxels, b->dim, b->n0, b->n1, b->n2, b->num, b->den,
                                                 b->rgb_acc, b->grid_pos, b->occupied_ids, b->first_key, b->counters + 3);
    e = cudaGetLastError();
  }
  const unsigned long long vv = static_cast<unsigned long long>(n_voxels);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b->counters, &vv, sizeof(vv), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n_bad, b->counters + 3, sizeof(n_bad), cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(df); cudaFree(dp); cudaFree(dw); cudaFree(dc);
  if (e != cudaSuccess) return cuda_fail(e, "builder import", __FILE__, __LINE__);
  b->id_upper = n_voxels;
  AVL_ARG(n_bad == 0, "grid_pos of the saved map lies outside this builder's grid");
  return AVL_OK;
}

int avl_builder_to_map(avl_builder* b, void* stream, avl_map** out) {
  AVL_ARG(b != nullptr && out != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t v = 0;
  int rc = read_counter(b, 0, &v, stream);
  if (rc) return rc;
  float* tmp = nullptr;
  AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&tmp), static_cast<size_t>(std::max<int64_t>(v, 1)) * b->dim * sizeof(float)));
  if (v > 0) export_feat_kernel<<<b->num_sms * 8, 256, 0, s>>>(b->num, b->den, v, b->dim, tmp);
  rc = avl_map_create(tmp, v, b->dim, AVL_ON_DEVICE, stream, out);
  cudaFree(tmp);
  return rc;
}

int avl_builder_export_keys(avl_builder* b, uint64_t* keys, int flags, void* stream) {
  AVL_ARG(b != nullptr && keys != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int64_t v = 0;
  int rc = read_counter(b, 0, &v, stream);
  if (rc) return rc;
  if (v == 0) return AVL_OK;
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(keys);
  if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dst), static_cast<size_t>(v) * sizeof(unsigned long long)));
  export_keys_kernel<<<b->num_sms * 8, 256, 0, s>>>(b->occupied_ids, b->first_key, b->cells, b->capacity, dst);
  cudaError_t e = cudaGetLastError();
  if (!(flags & AVL_ON_DEVICE)) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(keys, dst, static_cast<size_t>(v) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dst);
  }
  if (e != cudaSuccess) return cuda_fail(e, "export keys", __FILE__, __LINE__);
  return AVL_OK;
}

int avl_rank_keys(const uint64_t* keys_all, const int64_t* offsets, int32_t n_shards, int32_t shard,
                  int64_t* out_global_ids, int flags, void* stream) {
  AVL_ARG(offsets != nullptr, "offsets is NULL");
  AVL_ARG(n_shards >= 1 && n_shards <= kMaxShards && shard >= 0 && shard < n_shards, "invalid shard count");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ShardOffsets so;
  memset(&so, 0, sizeof(so));
  for (int i = 0; i <= n_shards; ++i) {
    so.off[i] = offsets[i];
    AVL_ARG(i == 0 ? offsets[0] == 0 : offsets[i] >= offsets[i - 1], "offsets must start at 0 and ascend");
  }
  const int64_t total = so.off[n_shards], n_own = so.off[shard + 1] - so.off[shard];
  if (n_own == 0) return AVL_OK;
  AVL_ARG(keys_all != nullptr && out_global_ids != nullptr, "NULL argument");
  const unsigned long long* k = reinterpret_cast<const unsigned long long*>(keys_all);
  int64_t* o = out_global_ids;
  unsigned long long* dk = nullptr;
  int64_t* dout = nullptr;
  if (!(flags & AVL_ON_DEVICE)) {
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dk), static_cast<size_t>(total) * sizeof(unsigned long long)));
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dout), static_cast<size_t>(n_own) * sizeof(int64_t));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dk, keys_all, static_cast<size_t>(total) * sizeof(unsigned long long), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { cudaFree(dk); cudaFree(dout); return cuda_fail(e, "rank keys staging", __FILE__, __LINE__); }
    k = dk;
    o = dout;
  }
  const int blocks = static_cast<int>(std::min<int64_t>((n_own + 255) / 256, 148 * 8));
  rank_keys_kernel<<<blocks, 256, 0, s>>>(k, so, n_shards, shard, o);
  cudaError_t e = cudaGetLastError();
  if (!(flags & AVL_ON_DEVICE)) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_global_ids, dout, static_cast<size_t>(n_own) * sizeof(int64_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dk);
    cudaFree(dout);
  }
  if (e != cudaSuccess) return cuda_fail(e, "rank keys", __FILE__, __LINE__);
  return AVL_OK;
}

int avl_bounds_create(avl_bounds** out) {
  AVL_ARG(out != nullptr, "NULL argument");
  *out = nullptr;
  int dev = 0;
  AVL_CUDA(cudaGetDevice(&dev));
  avl_bounds* b = new avl_bounds();
  cudaDeviceGetAttribute(&b->num_sms, cudaDevAttrMultiProcessorCount, dev);
  const double init[7] = {INFINITY, INFINITY, INFINITY, -INFINITY, -INFINITY, -INFINITY, 0.0};
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&b->acc), sizeof(init));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->partial), static_cast<size_t>(b->num_sms) * 8 * 7 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(b->acc, init, sizeof(init), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    avl_bounds_destroy(b);
    return cuda_fail(e, "bounds state", __FILE__, __LINE__);
  }
  *out = b;
  return AVL_OK;
}

int avl_bounds_destroy(avl_bounds* b) {
  if (!b) return AVL_OK;
  cudaFree(b->acc); cudaFree(b->partial); cudaFree(b->d_depth); cudaFree(b->d_sidx);
  delete b;
  return AVL_OK;
}

int avl_bounds_add_frame(avl_bounds* b, const avl_frame* f, int flags, void* stream) {
  AVL_ARG(b != nullptr && f != nullptr && f->depth != nullptr, "NULL argument");
  AVL_ARG(f->h >= 1 && f->w >= 1, "invalid frame shape");
  const int64_t npix = static_cast<int64_t>(f->h) * f->w;
  AVL_ARG(npix < (int64_t(1) << 31), "frame too large");
  const int32_t n_samples = f->sample_idx ? f->n_samples : static_cast<int32_t>(npix);
  AVL_ARG(n_samples >= 0 && n_samples <= npix, "n_samples out of range");
  if (n_samples == 0) return AVL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const float* depth = f->depth;
  const int32_t* sidx = f->sample_idx;
  int rc;
  if (!(flags & AVL_ON_DEVICE)) {
    const size_t depth_bytes = static_cast<size_t>(npix) * ((flags & AVL_DEPTH_U16_MM) ? sizeof(uint16_t) : sizeof(float));
    if ((rc = grow(&b->d_depth, &b->depth_elems, static_cast<size_t>(npix)))) return rc;
    AVL_CUDA(cudaMemcpyAsync(b->d_depth, f->depth, depth_bytes, cudaMemcpyHostToDevice, s));
    depth = b->d_depth;
    if (f->sample_idx) {
      if ((rc = grow(&b->d_sidx, &b->sidx_elems, static_cast<size_t>(n_samples)))) return rc;
      AVL_CUDA(cudaMemcpyAsync(b->d_sidx, f->sample_idx, static_cast<size_t>(n_samples) * sizeof(int32_t), cudaMemcpyHostToDevice, s));
      sidx = b->d_sidx;
    }
  }
  FrameGeom g;
  fill_geom(&g, f, flags);
  const int blocks = std::min((n_samples + 255) / 256, b->num_sms * 8);
  bounds_kernel<<<blocks, 256, 0, s>>>(g, depth, sidx, n_samples, b->partial);
  bounds_merge_kernel<<<1, 32, 0, s>>>(b->partial, blocks, b->acc);
  AVL_CUDA(cudaGetLastError());
  if (!(flags & AVL_ON_DEVICE)) AVL_CUDA(cudaStreamSynchronize(s));
  return AVL_OK;
}

int avl_bounds_get(avl_bounds* b, double pcd_min[3], double pcd_max[3], int64_t* n_points, void* stream) {
  AVL_ARG(b != nullptr && pcd_min != nullptr && pcd_max != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double acc[7];
  AVL_CUDA(cudaMemcpyAsync(acc, b->acc, sizeof(acc), cudaMemcpyDeviceToHost, s));
  AVL_CUDA(cudaStreamSynchronize(s));
  for (int c = 0; c < 3; ++c) { pcd_min[c] = acc[c]; pcd_max[c] = acc[3 + c]; }
  if (n_points) *n_points = static_cast<int64_t>(acc[6]);
  return AVL_OK;
}

}  // extern "C"
