// Spot detection on the scale-normalised LoG cube: threshold + 26/80-neighbour non-max
// suppression + compaction (skimage.feature.peak_local_max as blob_log calls it), ordering of
// the peaks and blob_log's _prune_blobs.  Call sites: modules/muscle_fish.py:333,661;
// modules/scope_utils3.py:403.
//
// nms_compact_kernel is a pure streaming kernel (R4 B/voxel): the threshold test rejects all but
// ~0.1 % of the voxels, and only the survivors fetch their neighbours (L2 hits).  Survivors are
// compacted with one warp-aggregated atomic per warp.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace mfish {

struct NmsArgs {
  const float* cube;  // [ns][nz][nx][ny]
  int ns, nz, nx, ny;
  float thr;
  int64_t* out_idx;
  float* out_val;
  uint32_t* count;
  int64_t capacity;
};

// Neighbour test for one super-threshold voxel.  Out-of-volume neighbours do not exist (the
// reference pads with 0, which can never exceed v > t >= 0).  Loads are issued in independent
// batches so that a test costs a few L2 round trips instead of 26 dependent ones:
//   1. the 4 in-plane / cross-plane face neighbours (x+-1, z+-1)   -> rejects most non-peaks
//   2. the remaining rows, 3 values per row, 9 rows in flight
__device__ __forceinline__ bool is_peak(const NmsArgs& a, int s, int z, int x, int y, float v) {
  const int64_t plane = (int64_t)a.nx * a.ny;
  const int64_t vol = plane * a.nz;
  const float* c = a.cube + s * vol + z * plane + (int64_t)x * a.ny + y;
  {
    const float xm = x > 0 ? __ldg(c - a.ny) : 0.f;
    const float xp = x + 1 < a.nx ? __ldg(c + a.ny) : 0.f;
    const float zm = z > 0 ? __ldg(c - plane) : 0.f;
    const float zp = z + 1 < a.nz ? __ldg(c + plane) : 0.f;
    if (fmaxf(fmaxf(xm, xp), fmaxf(zm, zp)) > v) return false;
  }
  const bool ym = y > 0, yp = y + 1 < a.ny;
  for (int ds = -1; ds <= 1; ++ds) {
    const int ss = s + ds;
    if (ss < 0 || ss >= a.ns) continue;
    float m = 0.f;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int zz = z + dz, xx = x + dx;
        if (zz >= 0 && zz < a.nz && xx >= 0 && xx < a.nx) {
          const float* row = c + ds * vol + dz * plane + dx * (int64_t)a.ny;
          const float r0 = ym ? __ldg(row - 1) : 0.f;
          const float r1 = __ldg(row);
          const float r2 = yp ? __ldg(row + 1) : 0.f;
          // the centre voxel itself takes part with its own value: harmless (v > v is false)
          m = fmaxf(m, fmaxf(r0, fmaxf(r1, r2)));
        }
      }
    }
    if (m > v) return false;
  }
  return true;
}

__device__ __forceinline__ void emit_peak(const NmsArgs& a, int s, int z, int x, int y, float v) {
  // warp-aggregated append
  unsigned active = __activemask();
  int leader = __ffs(active) - 1;
  int lane = threadIdx.x & 31;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(a.count, (uint32_t)__popc(active));
  base = __shfl_sync(active, base, leader);
  uint32_t pos = base + __popc(active & ((1u << lane) - 1));
  if ((int64_t)pos < a.capacity) {
    // reference raster index over the (x, y, z, s) cube
    a.out_idx[pos] = (((int64_t)x * a.ny + y) * a.nz + z) * a.ns + s;
    a.out_val[pos] = v;
  }
}

constexpr int NMS_UNROLL = 4;
constexpr int NMS_THREADS = 256;
constexpr int NMS_CHUNK = NMS_THREADS * NMS_UNROLL * 4;  // voxels scanned per block iteration

__device__ __forceinline__ void nms_decode(const NmsArgs& a, int64_t vox, int& s, int& z, int& x, int& y) {
  y = (int)(vox % a.ny);
  int64_t r = vox / a.ny;
  x = (int)(r % a.nx);
  r /= a.nx;
  z = (int)(r % a.nz);
  s = (int)(r / a.nz);
}

// Two phases per chunk so that no warp diverges on the (rare, expensive) neighbour test:
//   1. stream NMS_UNROLL x 128-bit loads per thread; a voxel that is above the threshold and not
//      beaten by its y neighbours inside the quad becomes a candidate (offset pushed to smem);
//   2. the block walks the candidate list with all lanes busy: 26/80-neighbour test, append.
__global__ void __launch_bounds__(NMS_THREADS) nms_compact_kernel(const NmsArgs a) {
  __shared__ uint16_t s_cand[NMS_CHUNK];
  __shared__ float s_val[NMS_CHUNK];
  __shared__ int s_n;
  const int64_t total_vox = (int64_t)a.ns * a.nz * a.nx * a.ny;
  const bool vec = ((a.ny & 3) == 0) && aligned16_dev(a.cube);
  const int64_t nchunks = (total_vox + NMS_CHUNK - 1) / NMS_CHUNK;
  const float thr = a.thr;
  for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int64_t v0 = ch * NMS_CHUNK;
    if (vec) {
      // ny % 4 == 0: every quad lies inside one row
      float4 q[NMS_UNROLL];
#pragma unroll
      for (int u = 0; u < NMS_UNROLL; ++u) {
        const int64_t off = v0 + (int64_t)(u * NMS_THREADS + threadIdx.x) * 4;
        q[u] = off < total_vox ? __ldg(reinterpret_cast<const float4*>(a.cube + off))
                               : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
#pragma unroll
      for (int u = 0; u < NMS_UNROLL; ++u) {
        const float v[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
        if (fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])) > thr) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float l = i > 0 ? v[i - 1] : -INFINITY;
            const float r = i < 3 ? v[i + 1] : -INFINITY;
            if (v[i] > thr && !(fmaxf(l, r) > v[i])) {
              const int pos = atomicAdd(&s_n, 1);
              s_cand[pos] = (uint16_t)((u * NMS_THREADS + threadIdx.x) * 4 + i);
              s_val[pos] = v[i];
            }
          }
        }
      }
    } else {
      for (int k = threadIdx.x; k < NMS_CHUNK; k += NMS_THREADS) {
        const int64_t off = v0 + k;
        if (off < total_vox) {
          const float v = __ldg(a.cube + off);
          if (v > thr) {
            const int pos = atomicAdd(&s_n, 1);
            s_cand[pos] = (uint16_t)k;
            s_val[pos] = v;
          }
        }
      }
    }
    __syncthreads();
    const int n = s_n;
    for (int k = threadIdx.x; k < n; k += NMS_THREADS) {
      int s, z, x, y;
      nms_decode(a, v0 + s_cand[k], s, z, x, y);
      const float v = s_val[k];
      if (is_peak(a, s, z, x, y, v)) emit_peak(a, s, z, x, y, v);
    }
    __syncthreads();
  }
}

// Second half of the fused LoG + NMS path: full neighbourhood test of the candidates the x pass emitted.
__global__ void __launch_bounds__(256) nms_verify_kernel(const NmsArgs a, const uint32_t* __restrict__ cand,
                                                         const uint32_t* __restrict__ cand_count, uint32_t cand_cap) {
  const uint32_t n = *cand_count;
  if (n > cand_cap) {  // candidate overflow: tell the host to run the stand-alone kernel
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.count = 0xFFFFFFFFu;
    return;
  }
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    int s, z, x, y;
    nms_decode(a, (int64_t)cand[k], s, z, x, y);
    const float v = __ldg(a.cube + cand[k]);
    if (is_peak(a, s, z, x, y, v)) emit_peak(a, s, z, x, y, v);
  }
}

int nms_verify_candidates(const float* cube, int64_t nz, int64_t nx, int64_t ny, float thr, const uint32_t* cand,
                          const uint32_t* cand_count, uint32_t cand_cap, int64_t* out_idx, float* out_val,
                          uint32_t* count, int64_t capacity, cudaStream_t st) {
  if (cudaMemsetAsync(count, 0, sizeof(uint32_t), st) != cudaSuccess) {
    set_error("nms_verify: memset failed");
    return MFISH_ERR_CUDA;
  }
  NmsArgs a{cube, 1, (int)nz, (int)nx, (int)ny, thr, out_idx, out_val, count, capacity};
  ProfScope prof("nms_verify", st);
  nms_verify_kernel<<<num_sms() * 4, 256, 0, st>>>(a, cand, cand_count, cand_cap);
  count_launch(1);
  return check_launch("nms_verify");
}

// ------------------------------------------------------------------------------------------
// ranking + equal-sigma prune
// ------------------------------------------------------------------------------------------
__global__ void iota_kernel(int32_t* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)i;
}
// rank r <- position n-1-r of the ascending raster order (reverse raster order)
__global__ void reverse_iota_kernel(int32_t* p, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int32_t)(n - 1 - i);
}

// descending value  <=>  ascending ~key
__global__ void desc_key_kernel(const float* val, uint32_t* key, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) key[i] = ~f32_to_key(val[i]);
}

__global__ void apply_rank_kernel(const int32_t* pos_of_rank, const int64_t* idxA, const float* valA,
                                  int64_t* idx_out, float* val_out, int32_t* rank_of_pos, int64_t n) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) {
    int32_t p = pos_of_rank[r];
    idx_out[r] = idxA[p];
    val_out[r] = valA[p];
    rank_of_pos[p] = (int32_t)r;
  }
}

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t* a, int64_t n, int64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct PruneArgs {
  const int64_t* idx_rank;  // by rank
  const int64_t* idxA;      // ascending
  const int32_t* rank_of_pos;
  int64_t n;
  int ns, nz, nx, ny;
  int cutoff_sq;
  int rad;
  uint8_t* alive;
};

// Equal sigma: peak r dies iff a peak with rank > r lies within squared distance cutoff_sq
// (sequential _prune_blobs in sorted pair order kills the lower index of every overlapping
// pair, and with equal sigma a peak can only be killed while it is the lower index).
__global__ void __launch_bounds__(128) prune_equal_kernel(const PruneArgs a) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n) return;
  int64_t id = a.idx_rank[r];
  const int s = (int)(id % a.ns);
  int64_t t = id / a.ns;
  const int z = (int)(t % a.nz);
  t /= a.nz;
  const int y = (int)(t % a.ny);
  const int x = (int)(t / a.ny);
  uint8_t alive = 1;
  for (int dx = -a.rad; dx <= a.rad && alive; ++dx) {
    int xx = x + dx;
    if (xx < 0 || xx >= a.nx) continue;
    for (int dy = -a.rad; dy <= a.rad && alive; ++dy) {
      int yy = y + dy;
      if (yy < 0 || yy >= a.ny) continue;
      int rem = a.cutoff_sq - dx * dx - dy * dy;
      if (rem < 0) continue;
      // contiguous z-run [z-dzmax, z+dzmax] maps to a contiguous raster range (step ns)
      int dzmax = (int)floorf(sqrtf((float)rem));
      while ((dzmax + 1) * (dzmax + 1) <= rem) ++dzmax;
      while (dzmax * dzmax > rem) --dzmax;
      int zlo = max(z - dzmax, 0), zhi = min(z + dzmax, a.nz - 1);
      int64_t klo = (((int64_t)xx * a.ny + yy) * a.nz + zlo) * a.ns;
      int64_t khi = (((int64_t)xx * a.ny + yy) * a.nz + zhi) * a.ns + (a.ns - 1);
      int64_t p = lower_bound_i64(a.idxA, a.n, klo);
      for (; p < a.n && a.idxA[p] <= khi; ++p) {
        int64_t j = a.idxA[p];
        if (j == id) continue;
        if ((int)(j % a.ns) != s) continue;  // equal-sigma rule applies within a scale
        if (a.rank_of_pos[p] > r) { alive = 0; break; }
      }
    }
  }
  a.alive[r] = alive;
}

// ------------------------------------------------------------------------------------------
// unequal-sigma candidate pairs (multi-scale blob_log)
// ------------------------------------------------------------------------------------------
struct PairArgs {
  const int64_t* idx_rank;
  int64_t n;
  int ns, nz, nx, ny;
  const int64_t* cutoff;  // device [ns*ns]
  int rad;                // max search radius in voxels
  const int64_t* idxA;    // ascending raster index
  const int32_t* rank_of_pos;
  int32_t* pairs;
  uint32_t* count;
  int64_t capacity;
};

__global__ void __launch_bounds__(128) overlap_pairs_kernel(const PairArgs a) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.n) return;
  int64_t id = a.idx_rank[r];
  const int s = (int)(id % a.ns);
  int64_t t = id / a.ns;
  const int z = (int)(t % a.nz);
  t /= a.nz;
  const int y = (int)(t % a.ny);
  const int x = (int)(t / a.ny);
  for (int dx = -a.rad; dx <= a.rad; ++dx) {
    int xx = x + dx;
    if (xx < 0 || xx >= a.nx) continue;
    for (int dy = -a.rad; dy <= a.rad; ++dy) {
      int yy = y + dy;
      if (yy < 0 || yy >= a.ny) continue;
      int zlo = max(z - a.rad, 0), zhi = min(z + a.rad, a.nz - 1);
      int64_t klo = (((int64_t)xx * a.ny + yy) * a.nz + zlo) * a.ns;
      int64_t khi = (((int64_t)xx * a.ny + yy) * a.nz + zhi) * a.ns + (a.ns - 1);
      int64_t p = lower_bound_i64(a.idxA, a.n, klo);
      for (; p < a.n && a.idxA[p] <= khi; ++p) {
        int32_t r2 = a.rank_of_pos[p];
        if (r2 <= r) continue;  // emit each pair once, i < j
        int64_t j = a.idxA[p];
        int s2 = (int)(j % a.ns);
        int zz = (int)((j / a.ns) % a.nz);
        int64_t d2 = (int64_t)dx * dx + (int64_t)dy * dy + (int64_t)(zz - z) * (zz - z);
        if (d2 <= a.cutoff[s * a.ns + s2]) {
          uint32_t pos = atomicAdd(a.count, 1u);
          if ((int64_t)pos < a.capacity) {
            a.pairs[2 * (int64_t)pos] = (int32_t)r;
            a.pairs[2 * (int64_t)pos + 1] = r2;
          }
        }
      }
    }
  }
}

// Shared front end: sort by raster index, then stable sort by descending value.
struct RankBuffers {
  int64_t* idxA = nullptr;
  float* valA = nullptr;
  int32_t* posA = nullptr;
  int32_t* pos_of_rank = nullptr;
  uint32_t* key = nullptr;
  uint32_t* key_sorted = nullptr;
  int32_t* rank_of_pos = nullptr;
  void* cub_tmp = nullptr;
};

static int rank_spots(const int64_t* idx, const float* val, int64_t n, int64_t* idx_out, float* val_out,
                      RankBuffers& b, int order, cudaStream_t st) {
  MFISH_REQUIRE(n < (1ll << 31), "rank: too many peaks");
  const int ni = (int)n;
  MFISH_CUDA(cudaMallocAsync((void**)&b.idxA, n * sizeof(int64_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&b.valA, n * sizeof(float), st));
  MFISH_CUDA(cudaMallocAsync((void**)&b.posA, n * sizeof(int32_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&b.pos_of_rank, n * sizeof(int32_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&b.key, n * sizeof(uint32_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&b.key_sorted, n * sizeof(uint32_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&b.rank_of_pos, n * sizeof(int32_t), st));
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t1, idx, b.idxA, val, b.valA, ni, 0, 64, st);
  cub::DeviceRadixSort::SortPairs(nullptr, t2, b.key, b.key_sorted, b.posA, b.pos_of_rank, ni, 0, 32, st);
  size_t tb = std::max(t1, t2);
  MFISH_CUDA(cudaMallocAsync(&b.cub_tmp, tb, st));
  MFISH_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, tb, idx, b.idxA, val, b.valA, ni, 0, 64, st));
  int blocks = (int)((n + 255) / 256);
  if (order == MFISH_ORDER_REVERSE_RASTER) {
    reverse_iota_kernel<<<blocks, 256, 0, st>>>(b.pos_of_rank, n);
  } else {
    iota_kernel<<<blocks, 256, 0, st>>>(b.posA, n);
    desc_key_kernel<<<blocks, 256, 0, st>>>(b.valA, b.key, n);
    MFISH_CUDA(cub::DeviceRadixSort::SortPairs(b.cub_tmp, tb, b.key, b.key_sorted, b.posA, b.pos_of_rank, ni,
                                               0, 32, st));
  }
  apply_rank_kernel<<<blocks, 256, 0, st>>>(b.pos_of_rank, b.idxA, b.valA, idx_out, val_out, b.rank_of_pos, n);
  count_launch(5);
  return check_launch("rank_spots");
}

static void free_rank(RankBuffers& b, cudaStream_t st) {
  void* ps[] = {b.idxA, b.valA, b.posA, b.pos_of_rank, b.key, b.key_sorted, b.rank_of_pos, b.cub_tmp};
  for (void* p : ps)
    if (p) cudaFreeAsync(p, st);
}

}  // namespace mfish

using namespace mfish;

extern "C" int mfish_nms_compact_f32(const float* cube_dev, int64_t ns, int64_t nz, int64_t nx, int64_t ny,
                                     float threshold, int64_t* out_idx_dev, float* out_val_dev,
                                     uint32_t* count_dev, int64_t capacity, void* stream) {
  MFISH_REQUIRE(cube_dev && out_idx_dev && out_val_dev && count_dev, "nms: null pointer");
  MFISH_REQUIRE(ns > 0 && nz > 0 && nx > 0 && ny > 0 && capacity > 0, "nms: bad sizes");
  MFISH_REQUIRE(ns < 65536 && nz < (1 << 30) && nx < (1 << 30) && ny < (1 << 30), "nms: dimension too large");
  MFISH_REQUIRE(threshold >= 0.f, "nms: threshold must be >= 0 (zero padding of the reference would differ)");
  cudaStream_t st = as_stream(stream);
  MFISH_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), st));
  NmsArgs a{cube_dev, (int)ns, (int)nz, (int)nx, (int)ny, threshold, out_idx_dev, out_val_dev, count_dev, capacity};
  int64_t nchunks = (ns * nz * nx * ny + NMS_CHUNK - 1) / NMS_CHUNK;
  int blocks = (int)std::min<int64_t>(nchunks, (int64_t)num_sms() * 32);
  ProfScope prof("nms_compact", st);
  nms_compact_kernel<<<blocks, NMS_THREADS, 0, st>>>(a);
  count_launch(1);
  return check_launch("nms_compact");
}

extern "C" int mfish_rank_prune(const int64_t* idx_dev, const float* val_dev, int64_t n, int64_t ns,
                                int64_t nz, int64_t nx, int64_t ny, int64_t cutoff_sq, int order,
                                int64_t* idx_out_dev, float* val_out_dev, uint8_t* alive_out_dev, void* stream) {
  MFISH_REQUIRE(n >= 0, "rank_prune: negative n");
  MFISH_REQUIRE(order == MFISH_ORDER_RESPONSE || order == MFISH_ORDER_REVERSE_RASTER, "rank_prune: bad order");
  if (n == 0) return MFISH_OK;
  MFISH_REQUIRE(idx_dev && val_dev && idx_out_dev && val_out_dev && alive_out_dev, "rank_prune: null pointer");
  MFISH_REQUIRE(cutoff_sq < (1 << 20), "rank_prune: cutoff too large");
  cudaStream_t st = as_stream(stream);
  ProfScope prof("rank_prune", st);
  RankBuffers b;
  int rc = rank_spots(idx_dev, val_dev, n, idx_out_dev, val_out_dev, b, order, st);
  if (rc == MFISH_OK) {
    if (cutoff_sq < 1) {
      MFISH_CUDA(cudaMemsetAsync(alive_out_dev, 1, n, st));
    } else {
      int rad = (int)floor(sqrt((double)cutoff_sq));
      PruneArgs a{idx_out_dev, b.idxA, b.rank_of_pos, n, (int)ns, (int)nz, (int)nx, (int)ny,
                  (int)cutoff_sq, rad, alive_out_dev};
      prune_equal_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a);
      count_launch(1);
      rc = check_launch("prune_equal");
    }
  }
  free_rank(b, st);
  return rc;
}

extern "C" int mfish_overlap_pairs(const int64_t* idx_ranked_dev, int64_t n, int64_t ns, int64_t nz,
                                   int64_t nx, int64_t ny, const int64_t* cutoff_sq_host, int32_t* pairs_dev,
                                   uint32_t* count_dev, int64_t capacity, void* stream) {
  MFISH_REQUIRE(n >= 0 && ns > 0 && ns <= 64, "overlap_pairs: bad sizes");
  MFISH_REQUIRE(count_dev != nullptr, "overlap_pairs: null count");
  cudaStream_t st = as_stream(stream);
  MFISH_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), st));
  if (n == 0) return MFISH_OK;
  MFISH_REQUIRE(idx_ranked_dev && cutoff_sq_host && pairs_dev, "overlap_pairs: null pointer");
  MFISH_REQUIRE(n < (1ll << 31), "overlap_pairs: too many peaks");
  int64_t cmax = -1;
  for (int i = 0; i < ns * ns; ++i) cmax = std::max(cmax, cutoff_sq_host[i]);
  if (cmax < 1) return MFISH_OK;
  int rad = (int)floor(sqrt((double)cmax));
  // ascending raster order + rank lookup (input is already in rank order)
  int64_t* idxA = nullptr;
  int32_t *rankv = nullptr, *rank_of_pos = nullptr;
  int64_t* cut_dev = nullptr;
  void* tmp = nullptr;
  MFISH_CUDA(cudaMallocAsync((void**)&idxA, n * sizeof(int64_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&rankv, n * sizeof(int32_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&rank_of_pos, n * sizeof(int32_t), st));
  MFISH_CUDA(cudaMallocAsync((void**)&cut_dev, ns * ns * sizeof(int64_t), st));
  MFISH_CUDA(cudaMemcpyAsync(cut_dev, cutoff_sq_host, ns * ns * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rankv, n);
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, idx_ranked_dev, idxA, rankv, rank_of_pos, (int)n, 0, 64, st);
  MFISH_CUDA(cudaMallocAsync(&tmp, tb, st));
  MFISH_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, idx_ranked_dev, idxA, rankv, rank_of_pos, (int)n, 0, 64, st));
  PairArgs a{idx_ranked_dev, n, (int)ns, (int)nz, (int)nx, (int)ny, cut_dev, rad, idxA, rank_of_pos,
             pairs_dev, count_dev, capacity};
  overlap_pairs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a);
  count_launch(3);
  int rc = check_launch("overlap_pairs");
  cudaFreeAsync(idxA, st);
  cudaFreeAsync(rankv, st);
  cudaFreeAsync(rank_of_pos, st);
  cudaFreeAsync(cut_dev, st);
  cudaFreeAsync(tmp, st);
  return rc;
}
