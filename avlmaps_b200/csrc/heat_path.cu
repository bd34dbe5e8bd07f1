// Object heat map: distance-decay to the nearest voxel of the indexed category.
//
// Replaces get_heatmap_from_mask_3d (reference avlmaps/utils/visualize_utils.py:29-49) and its twin
// HabitatLanguageRobot.get_vl_distribution_map_3d (avlmaps/robot/habitat_lang_robot.py:242-265): for
// every non-target voxel an O(N_target) numpy expression inside a Python loop -- the real wall-clock
// cost of AVLMap.index_object (avlmaps/map/avlmap.py:67-76).
//
// Exact: squared grid distances are integers, the minimum is taken on integers, and the float64
// tail (sqrt, / cell_size, * decay_rate, 1 - x, clip, cast) uses the reference's operation order with
// round-to-nearest intrinsics, so the result has the reference's bits.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "avl_internal.h"

namespace avl {
namespace {

constexpr int kHeatTile = 2048;

__global__ void __launch_bounds__(256)
compact_targets_kernel(const int32_t* __restrict__ pos, const uint8_t* __restrict__ mask, int64_t n,
                       int4* __restrict__ targets, uint32_t* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  for (int64_t base = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) - lane; base < n;
       base += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t i = base + lane;
    const bool t = i < n && mask[i] != 0;
    const uint32_t b = __ballot_sync(0xffffffffu, t);
    uint32_t off = 0;
    if (lane == 0 && b) off = atomicAdd(count, static_cast<uint32_t>(__popc(b)));
    off = __shfl_sync(0xffffffffu, off, 0);
    if (t) targets[off + __popc(b & ((1u << lane) - 1u))] = make_int4(pos[i * 3], pos[i * 3 + 1], pos[i * 3 + 2], 0);
  }
}

__global__ void __launch_bounds__(256)
heat_kernel(const int32_t* __restrict__ pos, const uint8_t* __restrict__ mask, int64_t n,
            const int4* __restrict__ targets, const uint32_t* __restrict__ count, double cell_size,
            double decay_rate, float* __restrict__ heat) {
  __shared__ int4 tile[kHeatTile];
  const uint32_t nt = *count;
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = i < n;
  int x = 0, y = 0, z = 0;
  bool is_target = false;
  if (active) {
    x = pos[i * 3]; y = pos[i * 3 + 1]; z = pos[i * 3 + 2];
    is_target = mask[i] != 0;
  }
  unsigned long long best = ~0ull;
  for (uint32_t t0 = 0; t0 < nt; t0 += kHeatTile) {
    const uint32_t m = min(static_cast<uint32_t>(kHeatTile), nt - t0);
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < m; j += blockDim.x) tile[j] = targets[t0 + j];
    __syncthreads();
    if (active && !is_target) {
#pragma unroll 4
      for (uint32_t j = 0; j < m; ++j) {
        const int4 t = tile[j];
        const long long dx = t.x - x, dy = t.y - y, dz = t.z - z;
        const unsigned long long d2 = static_cast<unsigned long long>(dx * dx + dy * dy + dz * dz);
        best = min(best, d2);
      }
    }
  }
  if (!active) return;
  if (is_target) {
    heat[i] = 1.0f;  // visualize_utils.py:45: heatmap initialised to ones, only non-targets overwritten
    return;
  }
  // visualize_utils.py:39-42: dist = norm / cell_size; sim = clip(1 - min_dist * decay_rate, 0, 1)
  const double dist = __ddiv_rn(__dsqrt_rn(static_cast<double>(best)), cell_size);
  double s = __dsub_rn(1.0, __dmul_rn(dist, decay_rate));
  s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
  heat[i] = static_cast<float>(s);
}


// ---- windowed search -------------------------------------------------------------------------
// The heat has finite support: beyond d = cell_size / decay_rate cells it is clipped to 0, so only targets
// inside that ball matter.  Targets are marked in a bitmap over their bounding box; every non-target voxel walks
// the ball's offsets in ascending squared distance and stops at the first marked cell -- its squared distance IS
// the minimum, an integer, so the float64 tail below produces the reference's bits.  No target inside the ball:
// the true distance exceeds the ball radius and the clip gives exactly 0.
struct HeatBox { int lo[3]; int dim[3]; };

__global__ void __launch_bounds__(256)
target_bbox_kernel(const int4* __restrict__ targets, const uint32_t* __restrict__ count, int* __restrict__ bbox) {
  const uint32_t nt = *count;
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) {
    const int4 t = targets[i];
    lo[0] = min(lo[0], t.x); lo[1] = min(lo[1], t.y); lo[2] = min(lo[2], t.z);
    hi[0] = max(hi[0], t.x); hi[1] = max(hi[1], t.y); hi[2] = max(hi[2], t.z);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    lo[c] = __reduce_min_sync(0xffffffffu, lo[c]);
    hi[c] = __reduce_max_sync(0xffffffffu, hi[c]);
  }
  if ((threadIdx.x & 31) == 0) {
    for (int c = 0; c < 3; ++c) { atomicMin(bbox + c, lo[c]); atomicMax(bbox + 3 + c, hi[c]); }
  }
}

__global__ void __launch_bounds__(256)
target_bitmap_kernel(const int4* __restrict__ targets, uint32_t nt, const HeatBox box, uint32_t* __restrict__ bits) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) {
    const int4 t = targets[i];
    const unsigned long long c = (static_cast<unsigned long long>(t.x - box.lo[0]) * box.dim[1] + (t.y - box.lo[1])) * box.dim[2] + (t.z - box.lo[2]);
    atomicOr(bits + (c >> 5), 1u << (c & 31));
  }
}

__global__ void __launch_bounds__(256)
heat_window_kernel(const int32_t* __restrict__ pos, const uint8_t* __restrict__ mask, int64_t n, const HeatBox box,
                   const uint32_t* __restrict__ bits, const int4* __restrict__ offs, int32_t n_offs, double cell_size,
                   double decay_rate, float* __restrict__ heat) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask[i] != 0) {
    heat[i] = 1.0f;
    return;
  }
  const int x = pos[i * 3] - box.lo[0], y = pos[i * 3 + 1] - box.lo[1], z = pos[i * 3 + 2] - box.lo[2];
  int best = -1;
  for (int k = 0; k < n_offs; ++k) {
    const int4 o = __ldg(offs + k);
    const int cx = x + o.x, cy = y + o.y, cz = z + o.z;
    if (static_cast<unsigned>(cx) < static_cast<unsigned>(box.dim[0]) && static_cast<unsigned>(cy) < static_cast<unsigned>(box.dim[1]) &&
        static_cast<unsigned>(cz) < static_cast<unsigned>(box.dim[2])) {
      const unsigned long long c = (static_cast<unsigned long long>(cx) * box.dim[1] + cy) * box.dim[2] + cz;
      if ((__ldg(bits + (c >> 5)) >> (c & 31)) & 1u) {
        best = o.w;
        break;
      }
    }
  }
  if (best < 0) {
    heat[i] = 0.0f;
    return;
  }
  const double dist = __ddiv_rn(__dsqrt_rn(static_cast<double>(best)), cell_size);
  double s = __dsub_rn(1.0, __dmul_rn(dist, decay_rate));
  s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
  heat[i] = static_cast<float>(s);
}

// ---- scatter from the targets -----------------------------------------------------------------
// When a dense uint32 grid over the targets' bounding box (grown by the support radius) fits, the roles flip:
// every target writes its squared distance into the cells of its ball (atomicMin), 10 k targets x 925 offsets
// instead of 1 M voxels x 925 offsets, and every voxel reads one cell.  Same minimum, same bits.
__global__ void __launch_bounds__(256)
heat_scatter_kernel(const int4* __restrict__ targets, uint32_t nt, const int4* __restrict__ offs, int32_t n_offs,
                    const HeatBox box, uint32_t* __restrict__ best) {
  const unsigned long long total = static_cast<unsigned long long>(nt) * n_offs;
  for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
    const int4 t = targets[i / n_offs];
    const int4 o = __ldg(offs + (i % n_offs));
    // box is the targets' bounding box grown by the radius on every side: every ball cell is inside
    const unsigned long long c = (static_cast<unsigned long long>(t.x + o.x - box.lo[0]) * box.dim[1] + (t.y + o.y - box.lo[1])) * box.dim[2] + (t.z + o.z - box.lo[2]);
    atomicMin(best + c, static_cast<uint32_t>(o.w));
  }
}

__global__ void __launch_bounds__(256)
heat_gather_kernel(const int32_t* __restrict__ pos, const uint8_t* __restrict__ mask, int64_t n, const HeatBox box,
                   const uint32_t* __restrict__ best, double cell_size, double decay_rate, float* __restrict__ heat) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (mask[i] != 0) {
    heat[i] = 1.0f;
    return;
  }
  const int x = pos[i * 3] - box.lo[0], y = pos[i * 3 + 1] - box.lo[1], z = pos[i * 3 + 2] - box.lo[2];
  uint32_t d2 = 0xFFFFFFFFu;
  if (static_cast<unsigned>(x) < static_cast<unsigned>(box.dim[0]) && static_cast<unsigned>(y) < static_cast<unsigned>(box.dim[1]) &&
      static_cast<unsigned>(z) < static_cast<unsigned>(box.dim[2]))
    d2 = __ldg(best + (static_cast<unsigned long long>(x) * box.dim[1] + y) * box.dim[2] + z);
  if (d2 == 0xFFFFFFFFu) {
    heat[i] = 0.0f;
    return;
  }
  const double dist = __ddiv_rn(__dsqrt_rn(static_cast<double>(d2)), cell_size);
  double s = __dsub_rn(1.0, __dmul_rn(dist, decay_rate));
  s = s < 0.0 ? 0.0 : (s > 1.0 ? 1.0 : s);
  heat[i] = static_cast<float>(s);
}

}  // namespace
}  // namespace avl

using namespace avl;

namespace {
// Grow-only device scratch of the heat entry points, one per host thread.  cudaMalloc / cudaFree per call cost
// milliseconds (cudaFree synchronises the device) -- measured: 9 ms of allocator time around 50 us of kernels.
struct HeatScratch {
  enum { kTargets, kCount, kBBox, kBits, kOffs, kPos, kMask, kHeat, kSlots };
  void* buf[kSlots] = {};
  size_t cap[kSlots] = {};
  int device = -1;
  cudaError_t get(int slot, size_t bytes, void** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev != device) {  // another device became current: drop the old buffers (no reuse across devices)
      release();
      device = dev;
    }
    if (cap[slot] < bytes) {
      if (buf[slot]) cudaFree(buf[slot]);
      buf[slot] = nullptr;
      cap[slot] = 0;
      const size_t want = bytes + bytes / 4 + 256;
      e = cudaMalloc(&buf[slot], want);
      if (e != cudaSuccess) return e;
      cap[slot] = want;
    }
    *out = buf[slot];
    return cudaSuccess;
  }
  void release() {
    for (int i = 0; i < kSlots; ++i) {
      if (buf[i]) cudaFree(buf[i]);
      buf[i] = nullptr;
      cap[i] = 0;
    }
  }
  // The pool is shared by consecutive calls of this thread.  A device-pointer call returns while its kernels may
  // still read the pool, so every call records `done` on its stream when it has enqueued its last kernel, and the next
  // call makes ITS stream wait for that event first: calls on different streams then use the pool one after the other.
  cudaEvent_t done = nullptr;
  bool pending = false;
  cudaError_t acquire(cudaStream_t s) {
    if (!done) {
      cudaError_t e = cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
    return pending ? cudaStreamWaitEvent(s, done, 0) : cudaSuccess;
  }
  void mark(cudaStream_t s) {
    if (done && cudaEventRecord(done, s) == cudaSuccess) pending = true;
  }
  ~HeatScratch() {}  // process teardown: the driver reclaims the memory; calling cudaFree there can race the runtime's own exit
};
thread_local HeatScratch g_heat_scratch;
}  // namespace

extern "C" int avl_heat_from_mask_3d(const int32_t* grid_pos, const uint8_t* mask, int64_t n, double cell_size,
                                     double decay_rate, float* out_heat, int flags, void* stream) {
  AVL_ARG(n >= 0 && n < (int64_t(1) << 31), "n out of range");
  AVL_ARG(n == 0 || (grid_pos && mask && out_heat), "NULL argument");
  AVL_ARG(cell_size > 0.0, "cell_size must be > 0");
  if (n == 0) return AVL_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int32_t* d_pos = nullptr;
  uint8_t* d_mask = nullptr;
  float* d_heat = nullptr;
  int4* d_targets = nullptr;
  uint32_t* d_count = nullptr;
  int* d_bbox = nullptr;
  uint32_t* d_bits = nullptr;
  int4* d_offs = nullptr;
  int rc = AVL_OK;
  cudaError_t e = cudaSuccess;
  const bool host = !(flags & AVL_ON_DEVICE);
  do {
    HeatScratch& hs = g_heat_scratch;
    e = hs.acquire(s);
    if (e == cudaSuccess) e = hs.get(HeatScratch::kTargets, static_cast<size_t>(n) * sizeof(int4), reinterpret_cast<void**>(&d_targets));
    if (e == cudaSuccess) e = hs.get(HeatScratch::kCount, sizeof(uint32_t), reinterpret_cast<void**>(&d_count));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_count, 0, sizeof(uint32_t), s);
    const int32_t* pos = grid_pos;
    const uint8_t* msk = mask;
    float* heat = out_heat;
    if (host && e == cudaSuccess) {
      e = hs.get(HeatScratch::kPos, static_cast<size_t>(n) * 3 * sizeof(int32_t), reinterpret_cast<void**>(&d_pos));
      if (e == cudaSuccess) e = hs.get(HeatScratch::kMask, static_cast<size_t>(n), reinterpret_cast<void**>(&d_mask));
      if (e == cudaSuccess) e = hs.get(HeatScratch::kHeat, static_cast<size_t>(n) * sizeof(float), reinterpret_cast<void**>(&d_heat));
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_pos, grid_pos, static_cast<size_t>(n) * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, s);
      if (e == cudaSuccess) e = cudaMemcpyAsync(d_mask, mask, static_cast<size_t>(n), cudaMemcpyHostToDevice, s);
      pos = d_pos; msk = d_mask; heat = d_heat;
    }
    if (e != cudaSuccess) break;
    const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
    compact_targets_kernel<<<std::min(blocks, 1184u), 256, 0, s>>>(pos, msk, n, d_targets, d_count);
    uint32_t nt = 0;
    e = cudaMemcpyAsync(&nt, d_count, sizeof(nt), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) break;
    if (nt == 0) {  // the reference's np.argmin raises on an empty target set
      set_error("heat_from_mask_3d: the mask selects no voxel");
      rc = AVL_ERR_ARG;
      break;
    }
    // window or brute force?  ball = offsets within the support radius; brute force costs nt distance evaluations per voxel
    bool windowed = false;
    const double support = cell_size / decay_rate;  // heat > 0 only below this many cells
    const bool force_brute = getenv("AVL_HEAT_BRUTE") != nullptr;  // A/B and tests of the brute-force kernel
    if (!force_brute && decay_rate > 0.0 && support < 40.0) {
      const int rmax = static_cast<int>(std::floor(support)) + 1;
      std::vector<int4> offs;
      for (int dx = -rmax; dx <= rmax; ++dx)
        for (int dy = -rmax; dy <= rmax; ++dy)
          for (int dz = -rmax; dz <= rmax; ++dz) {
            const int d2 = dx * dx + dy * dy + dz * dz;
            if (d2 <= rmax * rmax) offs.push_back(make_int4(dx, dy, dz, d2));
          }
      std::sort(offs.begin(), offs.end(), [](const int4& a, const int4& b) { return a.w < b.w; });
      // measured on B200 (1M voxels, 10 k targets): brute force 0.66 us per target, bitmap walk 2.8 us per ball
      // offset, scatter 0.11 us per ball offset (per 10 k targets)
      const bool scatter_pays = static_cast<double>(offs.size()) * 20.0 < static_cast<double>(n);
      const bool walk_pays = static_cast<double>(offs.size()) * 5.0 < static_cast<double>(nt);
      if (scatter_pays || walk_pays) {
        int h_bbox[6] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN};
        e = hs.get(HeatScratch::kBBox, sizeof(h_bbox), reinterpret_cast<void**>(&d_bbox));
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_bbox, h_bbox, sizeof(h_bbox), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) break;
        target_bbox_kernel<<<std::min((nt + 255u) / 256u, 592u), 256, 0, s>>>(d_targets, d_count, d_bbox);
        e = cudaMemcpyAsync(h_bbox, d_bbox, sizeof(h_bbox), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) break;
        HeatBox box;
        double cells = 1.0;
        for (int c = 0; c < 3; ++c) {
          box.lo[c] = h_bbox[c];
          box.dim[c] = h_bbox[3 + c] - h_bbox[c] + 1;
          cells *= static_cast<double>(box.dim[c]);
        }
        // preferred: scatter from the targets into a dense uint32 grid over the box grown by the radius
        double gcells = 1.0;
        HeatBox gbox;
        for (int c = 0; c < 3; ++c) {
          gbox.lo[c] = box.lo[c] - rmax;
          gbox.dim[c] = box.dim[c] + 2 * rmax;
          gcells *= static_cast<double>(gbox.dim[c]);
        }
        if (scatter_pays && gcells * 4.0 <= 2.0e9 && getenv("AVL_HEAT_BITMAP") == nullptr) {
          const size_t gwords = static_cast<size_t>(gcells);
          e = hs.get(HeatScratch::kBits, gwords * sizeof(uint32_t), reinterpret_cast<void**>(&d_bits));
          if (e == cudaSuccess) e = cudaMemsetAsync(d_bits, 0xFF, gwords * sizeof(uint32_t), s);
          if (e == cudaSuccess) e = hs.get(HeatScratch::kOffs, offs.size() * sizeof(int4), reinterpret_cast<void**>(&d_offs));
          if (e == cudaSuccess) e = cudaMemcpyAsync(d_offs, offs.data(), offs.size() * sizeof(int4), cudaMemcpyHostToDevice, s);
          if (e != cudaSuccess) break;
          const unsigned long long work = static_cast<unsigned long long>(nt) * offs.size();
          heat_scatter_kernel<<<static_cast<unsigned>(std::min<unsigned long long>((work + 255) / 256, 148ull * 32)), 256, 0, s>>>(
              d_targets, nt, d_offs, static_cast<int32_t>(offs.size()), gbox, d_bits);
          heat_gather_kernel<<<blocks, 256, 0, s>>>(pos, msk, n, gbox, d_bits, cell_size, decay_rate, heat);
          e = cudaGetLastError();
          if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // offs (host vector) is released below
          windowed = true;
        } else if ((walk_pays || getenv("AVL_HEAT_BITMAP") != nullptr) && cells <= 8.0e9) {  // bitmap of the targets + per-voxel walk of the ball
          const size_t words = static_cast<size_t>((static_cast<unsigned long long>(cells) + 31) / 32);
          e = hs.get(HeatScratch::kBits, words * sizeof(uint32_t), reinterpret_cast<void**>(&d_bits));
          if (e == cudaSuccess) e = cudaMemsetAsync(d_bits, 0, words * sizeof(uint32_t), s);
          if (e == cudaSuccess) e = hs.get(HeatScratch::kOffs, offs.size() * sizeof(int4), reinterpret_cast<void**>(&d_offs));
          if (e == cudaSuccess) e = cudaMemcpyAsync(d_offs, offs.data(), offs.size() * sizeof(int4), cudaMemcpyHostToDevice, s);
          if (e != cudaSuccess) break;
          target_bitmap_kernel<<<std::min((nt + 255u) / 256u, 1184u), 256, 0, s>>>(d_targets, nt, box, d_bits);
          heat_window_kernel<<<blocks, 256, 0, s>>>(pos, msk, n, box, d_bits, d_offs, static_cast<int32_t>(offs.size()),
                                                    cell_size, decay_rate, heat);
          e = cudaGetLastError();
          if (e == cudaSuccess) e = cudaStreamSynchronize(s);  // offs (host vector) and the scratch are released below
          windowed = true;
        }
      }
    }
    if (!windowed) heat_kernel<<<blocks, 256, 0, s>>>(pos, msk, n, d_targets, d_count, cell_size, decay_rate, heat);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (host && e == cudaSuccess) e = cudaMemcpyAsync(out_heat, d_heat, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, s);
    hs.mark(s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  } while (0);
  if (e != cudaSuccess && rc == AVL_OK) rc = cuda_fail(e, "heat_from_mask_3d", __FILE__, __LINE__);
  return rc;  // the scratch stays in g_heat_scratch for the next call
}

// ------------------------------------------------------------------------------------------------
// 2-D heat from point sources: the per-frame / per-segment full-grid distance_transform_edt loops of
// AVLMap.index_area_2d (reference avlmaps/map/avlmap.py:78-98, combine = max, float64 result) and
// AVLMap.index_sound_2d (avlmap.py:111-133, combine = sum accumulated in float32, in segment order).
// The EDT of a map with a few marked cells is the distance to the nearest marked cell: one thread per
// grid cell, sources staged in shared memory, integer squared distances, float64 tail in the
// reference's operation order.
namespace avl {
namespace {

constexpr int kSrcTile = 1024;

// sources are sorted by group; group g owns sources [gstart[g], gstart[g+1])
__global__ void __launch_bounds__(256)
heat2d_kernel(const int32_t* __restrict__ cells, const int32_t* __restrict__ gstart, const float* __restrict__ conf,
              int32_t n_groups, int32_t rows, int32_t cols, double decay_rate, int32_t mode, void* __restrict__ out) {
  __shared__ int2 tile[kSrcTile];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool active = i < static_cast<int64_t>(rows) * cols;
  const int r = active ? static_cast<int>(i / cols) : 0, c = active ? static_cast<int>(i % cols) : 0;
  double acc_max = 0.0;  // dist_map starts at zeros (avlmap.py:79)
  float acc_sum = 0.f;   // avlmap.py:113
  for (int g = 0; g < n_groups; ++g) {
    const int s0 = gstart[g], s1 = gstart[g + 1];
    if (s1 <= s0) continue;  // frame outside the grid: `continue` (avlmap.py:88-89)
    unsigned long long best = ~0ull;
    for (int t0 = s0; t0 < s1; t0 += kSrcTile) {
      const int m = min(kSrcTile, s1 - t0);
      __syncthreads();
      for (int j = threadIdx.x; j < m; j += blockDim.x) tile[j] = make_int2(cells[2 * (t0 + j)], cells[2 * (t0 + j) + 1]);
      __syncthreads();
      if (active)
        for (int j = 0; j < m; ++j) {
          const long long dr = tile[j].x - r, dc = tile[j].y - c;
          best = min(best, static_cast<unsigned long long>(dr * dr + dc * dc));
        }
    }
    if (!active) continue;
    const double dist = __dsqrt_rn(static_cast<double>(best));
    const float con = conf[g];
    if (mode == 0) {
      // area: tmp = ones * s - dists * decay; clip(tmp, 0, 1); dist_map = max(dist_map, tmp)   (avlmap.py:93-97)
      // (a source with s == 0 leaves tmp_dist_map all zero; the reference's EDT then measures to a virtual
      //  background cell, but -dists * decay clips to 0 either way)
      double v = __dsub_rn(static_cast<double>(con), __dmul_rn(dist, decay_rate));
      v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
      acc_max = acc_max > v ? acc_max : v;
    } else {
      // sound: reduct = con * dists * decay; tmp = ones * con - reduct; tmp[tmp < 0] = 0; dist_map += tmp
      // (float32 accumulator, float64 addend: avlmap.py:124-131)
      const double reduct = __dmul_rn(__dmul_rn(static_cast<double>(con), dist), decay_rate);
      double v = __dsub_rn(static_cast<double>(con), reduct);
      if (v < 0.0) v = 0.0;
      acc_sum = static_cast<float>(__dadd_rn(static_cast<double>(acc_sum), v));
    }
  }
  if (!active) return;
  if (mode == 0) reinterpret_cast<double*>(out)[i] = acc_max;
  else reinterpret_cast<float*>(out)[i] = acc_sum;
}

}  // namespace
}  // namespace avl

extern "C" int avl_heat2d_sources(const int32_t* cells, const int32_t* group_start, const float* conf, int32_t n_groups,
                                  int32_t rows, int32_t cols, double decay_rate, int32_t mode, void* out_heat,
                                  int flags, void* stream) {
  AVL_ARG(rows >= 1 && cols >= 1 && n_groups >= 0 && (mode == 0 || mode == 1), "invalid argument");
  AVL_ARG(out_heat != nullptr && (n_groups == 0 || (group_start && conf)), "NULL argument");
  if (flags & AVL_ON_DEVICE) {
    set_error("avl_heat2d_sources takes host pointers (a handful of sources, one small grid)");
    return AVL_ERR_UNSUPPORTED;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = static_cast<int64_t>(rows) * cols;
  const size_t esz = mode == 0 ? sizeof(double) : sizeof(float);
  const int n_src = n_groups > 0 ? group_start[n_groups] : 0;
  AVL_ARG(n_src == 0 || cells != nullptr, "cells is NULL");
  int32_t *d_cells = nullptr, *d_gs = nullptr;
  float* d_conf = nullptr;
  void* d_out = nullptr;
  cudaError_t e = cudaMalloc(&d_out, n * esz);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_cells), std::max(1, n_src) * 2 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_gs), (n_groups + 1) * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_conf), std::max(1, n_groups) * sizeof(float));
  if (e == cudaSuccess && n_src) e = cudaMemcpyAsync(d_cells, cells, static_cast<size_t>(n_src) * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && n_groups) e = cudaMemcpyAsync(d_gs, group_start, (n_groups + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && n_groups) e = cudaMemcpyAsync(d_conf, conf, n_groups * sizeof(float), cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) {
    heat2d_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(d_cells, d_gs, d_conf, n_groups, rows, cols,
                                                                         decay_rate, mode, d_out);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_heat, d_out, n * esz, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_cells); cudaFree(d_gs); cudaFree(d_conf); cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "heat2d_sources", __FILE__, __LINE__);
  return AVL_OK;
}


// ---------------------------------------------------------------- image modality: planar distance decay
// AVLMap.index_image after localisation (avlmaps/map/avlmap.py:156-162):
//   dists = np.linalg.norm((grid_pos - [row, col, height])[:, :2], axis=1); sim = np.clip(con - decay * dists, 0, 1)
// float64 throughout; sqrt(dx*dx + dy*dy) with separately rounded operations like numpy's norm.
namespace avl {
namespace {
__global__ void __launch_bounds__(256)
heat_planar_kernel(const int32_t* __restrict__ pos, int64_t n, double row, double col, double con, double decay,
                   double* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const double dx = __dsub_rn(static_cast<double>(pos[i * 3]), row);
    const double dy = __dsub_rn(static_cast<double>(pos[i * 3 + 1]), col);
    const double d = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    const double v = __dsub_rn(con, __dmul_rn(decay, d));
    out[i] = fmin(fmax(v, 0.0), 1.0);
  }
}
}  // namespace
}  // namespace avl

extern "C" int avl_heat_planar(const int32_t* grid_pos, int64_t n, double row, double col, double con,
                               double decay_rate, double* out_heat, int flags, void* stream) {
  using namespace avl;
  AVL_ARG(n >= 0, "n < 0");
  if (n == 0) return AVL_OK;
  AVL_ARG(grid_pos != nullptr && out_heat != nullptr, "NULL argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int32_t* p = grid_pos;
  double* o = out_heat;
  int32_t* dp = nullptr;
  double* dout = nullptr;
  if (!(flags & AVL_ON_DEVICE)) {
    AVL_CUDA(cudaMalloc(reinterpret_cast<void**>(&dp), static_cast<size_t>(n) * 3 * sizeof(int32_t)));
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&dout), static_cast<size_t>(n) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dp, grid_pos, static_cast<size_t>(n) * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { cudaFree(dp); cudaFree(dout); return cuda_fail(e, "heat planar staging", __FILE__, __LINE__); }
    p = dp;
    o = dout;
  }
  const int blocks = static_cast<int>(std::min<int64_t>((n + 255) / 256, 148 * 8));
  heat_planar_kernel<<<blocks, 256, 0, s>>>(p, n, row, col, con, decay_rate, o);
  cudaError_t e = cudaGetLastError();
  if (!(flags & AVL_ON_DEVICE)) {
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_heat, dout, static_cast<size_t>(n) * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(dp);
    cudaFree(dout);
  }
  if (e != cudaSuccess) return cuda_fail(e, "heat planar", __FILE__, __LINE__);
  return AVL_OK;
}


// ---------------------------------------------------------------- 2-D heat: final min-max and the 2-D -> 3-D lift
// dist_map = (dist_map - min) / (max - min) in the map's own dtype (avlmap.py:97 float64, :131 float32), then
// heatmap_3d[id] = heatmap_2d[row, col] for every occupied cell (avlmap.py:100-109, 135-144) = a gather through
// grid_pos.  min / max are exact, the two IEEE operations are the reference's, so the bits match numpy's.
namespace avl {
namespace {
template <typename T>
__global__ void __launch_bounds__(1024)
minmax2d_kernel(const T* __restrict__ v, int64_t n, T* __restrict__ mm) {
  T lo = v[0], hi = v[0];
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const T x = v[i];
    lo = x < lo ? x : lo;
    hi = x > hi ? x : hi;
  }
  __shared__ T slo[32], shi[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w) { lo = slo[w] < lo ? slo[w] : lo; hi = shi[w] > hi ? shi[w] : hi; }
    mm[0] = lo;
    mm[1] = hi;
  }
}
__global__ void __launch_bounds__(256) normalize2d_f64_kernel(double* __restrict__ v, int64_t n, const double* __restrict__ mm) {
  const double lo = mm[0], range = __dsub_rn(mm[1], mm[0]);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    v[i] = __ddiv_rn(__dsub_rn(v[i], lo), range);
}
__global__ void __launch_bounds__(256) normalize2d_f32_kernel(float* __restrict__ v, int64_t n, const float* __restrict__ mm) {
  const float lo = mm[0], range = __fsub_rn(mm[1], mm[0]);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    v[i] = __fdiv_rn(__fsub_rn(v[i], lo), range);
}
template <typename T>
__global__ void __launch_bounds__(256)
lift2d_kernel(const T* __restrict__ heat2d, int32_t rows, int32_t cols, const int32_t* __restrict__ pos, int64_t n,
              float* __restrict__ out) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int r = pos[i * 3], c = pos[i * 3 + 1];
    if (r < 0) r += rows;  // multi-floor maps keep unwrapped (negative) indices in grid_pos
    if (c < 0) c += cols;
    out[i] = (r >= 0 && r < rows && c >= 0 && c < cols) ? static_cast<float>(heat2d[static_cast<int64_t>(r) * cols + c]) : 0.f;
  }
}
}  // namespace
}  // namespace avl

extern "C" int avl_heat2d_normalize_lift(void* heat2d, int32_t is_f64, int32_t rows, int32_t cols, int32_t normalize,
                                         const int32_t* grid_pos, int64_t n, float* out_heat3d, int flags, void* stream) {
  using namespace avl;
  AVL_ARG(heat2d != nullptr && rows >= 1 && cols >= 1, "invalid 2-D heat map");
  AVL_ARG(n >= 0 && (n == 0 || (grid_pos != nullptr && out_heat3d != nullptr)), "grid_pos / out_heat3d is NULL");
  AVL_ARG(!(flags & AVL_ON_DEVICE), "host pointers only");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t cells = static_cast<int64_t>(rows) * cols;
  const size_t esz = is_f64 ? sizeof(double) : sizeof(float);
  HeatScratch& hs = g_heat_scratch;
  void *d2 = nullptr, *dmm = nullptr;
  int32_t* dpos = nullptr;
  float* d3 = nullptr;
  cudaError_t e = hs.get(HeatScratch::kBits, cells * esz, &d2);
  if (e == cudaSuccess) e = hs.get(HeatScratch::kBBox, 2 * sizeof(double), &dmm);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d2, heat2d, cells * esz, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess && normalize) {
    if (is_f64) {
      minmax2d_kernel<double><<<1, 1024, 0, s>>>(static_cast<const double*>(d2), cells, static_cast<double*>(dmm));
      normalize2d_f64_kernel<<<592, 256, 0, s>>>(static_cast<double*>(d2), cells, static_cast<const double*>(dmm));
    } else {
      minmax2d_kernel<float><<<1, 1024, 0, s>>>(static_cast<const float*>(d2), cells, static_cast<float*>(dmm));
      normalize2d_f32_kernel<<<592, 256, 0, s>>>(static_cast<float*>(d2), cells, static_cast<const float*>(dmm));
    }
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(heat2d, d2, cells * esz, cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess && n > 0) {
    e = hs.get(HeatScratch::kPos, static_cast<size_t>(n) * 3 * sizeof(int32_t), reinterpret_cast<void**>(&dpos));
    if (e == cudaSuccess) e = hs.get(HeatScratch::kHeat, static_cast<size_t>(n) * sizeof(float), reinterpret_cast<void**>(&d3));
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpos, grid_pos, static_cast<size_t>(n) * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
      const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((n + 255) / 256, 1184));
      if (is_f64) lift2d_kernel<double><<<blocks, 256, 0, s>>>(static_cast<const double*>(d2), rows, cols, dpos, n, d3);
      else lift2d_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float*>(d2), rows, cols, dpos, n, d3);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_heat3d, d3, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, s);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) return cuda_fail(e, "heat2d_normalize_lift", __FILE__, __LINE__);
  return AVL_OK;
}
