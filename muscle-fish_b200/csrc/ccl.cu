// Connected-component labelling of a binary volume (the part of threshold_fiber /
// threshold_nuclei that round 1 left on the host): skimage.morphology.label / remove_small_objects /
// remove_small_holes as modules/muscle_fish.py:182-184, 266-272 call them.
//
// The volume is taken in the REFERENCE's own (x, y, z) C order (z fastest), so that a voxel's linear index is the
// raster index scipy / skimage scan in: a component's label is then simply the rank of its smallest index.
// Union-find in place over an int32 parent volume (Playne & Hawick style): rows (the fastest axis) are linked by
// the initialisation (every voxel points at the start of its run), then every foreground voxel links itself to its
// foreground neighbours in the previous rows (2 of the 6 face neighbours -- skimage connectivity 1 -- or 12 of
// the 26 -- connectivity 3, skimage's default for label()) with atomicMin on the larger root, skipping pairs that
// the pair to their left already joins; finds halve their paths; a flatten pass makes every voxel point at its
// component's smallest index.  Ranking the roots (a scan) and gathering the rank
// per voxel are left to the caller's tensor library: they are plain streaming passes.
#include "common.cuh"

namespace mfish {

__device__ __forceinline__ int ccl_find(const int* parent, int i) {
  int p = parent[i];
  while (p != i) {
    i = p;
    p = parent[i];
  }
  return i;
}
// find with path halving; the shortcut is written with atomicMin, so a link only ever moves towards smaller
// indices (an ancestor), whatever races with it
__device__ __forceinline__ int ccl_find_compress(int* parent, int i) {
  int p = parent[i];
  while (p != i) {
    const int g = parent[p];
    if (g != p) atomicMin(&parent[i], g);
    i = p;
    p = g;
  }
  return i;
}

__device__ __forceinline__ void ccl_union(int* parent, int a, int b) {
  while (true) {
    a = ccl_find_compress(parent, a);
    b = ccl_find_compress(parent, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(&parent[a], b);  // a > b: hang the larger root under the smaller
    if (old == a) return;
    a = old;  // someone re-rooted a in the meantime: go on from there
  }
}

// One thread per row (the fastest axis): every foreground voxel starts out pointing at the first voxel of its
// RUN along the row, so the row direction needs no union at all and chains start one link deep.
__global__ void __launch_bounds__(128) ccl_init_rows_kernel(const uint8_t* __restrict__ mask, int* __restrict__ parent,
                                                            int64_t nrows, int nc) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const int64_t base = row * nc;
  int start = -1;
  for (int c = 0; c < nc; ++c) {
    const bool fg = mask[base + c] != 0;
    if (fg && start < 0) start = (int)(base + c);
    if (!fg) start = -1;
    parent[base + c] = start;
  }
}

// dims (na, nb, nc), c fastest.  FULL: 26-neighbourhood (13 predecessors), else the 6-neighbourhood (3).
template <bool FULL>
__global__ void __launch_bounds__(256) ccl_merge_kernel(int* parent, int na, int nb, int nc) {
  const int64_t n = (int64_t)na * nb * nc;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (parent[i] < 0) continue;
    const int c = (int)(i % nc);
    const int64_t r = i / nc;
    const int b = (int)(r % nb);
    const int a = (int)(r / nb);
    // (the row direction is linked by the initialisation)  A neighbour in another row needs a union only where
    // the pair to the LEFT of it is not a foreground pair as well -- that pair's union already joins the two runs
    const bool left = c > 0 && parent[i - 1] >= 0;
    if (!FULL) {
      if (b > 0 && parent[i - nc] >= 0 && !(left && parent[i - nc - 1] >= 0))
        ccl_union(parent, (int)i, (int)(i - nc));
      const int64_t sa = (int64_t)nb * nc;
      if (a > 0 && parent[i - sa] >= 0 && !(left && parent[i - sa - 1] >= 0)) ccl_union(parent, (int)i, (int)(i - sa));
    } else {
      // predecessors in raster order: (da, db, dc) lexicographically below (0, 0, 0)
#pragma unroll
      for (int da = -1; da <= 0; ++da) {
#pragma unroll
        for (int db = -1; db <= 1; ++db) {
#pragma unroll
          for (int dc = -1; dc <= 1; ++dc) {
            if (da == 0 && db >= 0) continue;  // same row: linked by the initialisation; later rows: their turn
            const int a2 = a + da, b2 = b + db, c2 = c + dc;
            if (a2 < 0 || b2 < 0 || b2 >= nb || c2 < 0 || c2 >= nc) continue;
            const int64_t j = ((int64_t)a2 * nb + b2) * nc + c2;
            if (parent[j] < 0) continue;
            // j's run already meets i's run through the pair one step to the left (i-1 with j-1, j or j+1 all lie
            // in the same 3-wide window): skip when i-1 is foreground and the window entry left of j is too
            if (left && dc >= 0 && parent[j - 1] >= 0) continue;
            ccl_union(parent, (int)i, (int)j);
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) ccl_flatten_kernel(int* parent, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (parent[i] < 0) continue;
    parent[i] = ccl_find(parent, (int)i);
  }
}

// voxel count per value of an int32 volume (np.bincount): every thread walks CCL_RUN consecutive voxels and adds
// one count per RUN of equal values -- a volume that is one giant component (the complement of a sparse mask, what
// remove_small_holes labels) would otherwise send every voxel's atomic to the same address
constexpr int CCL_RUN = 64;
__global__ void __launch_bounds__(256) label_sizes_kernel(const int* __restrict__ labels, int64_t n, int nvalues,
                                                          unsigned long long* __restrict__ sizes) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = t * CCL_RUN;
  if (i0 >= n) return;
  const int64_t i1 = min(n, i0 + CCL_RUN);
  int cur = labels[i0];
  unsigned cnt = 1;
  for (int64_t i = i0 + 1; i < i1; ++i) {
    const int v = labels[i];
    if (v == cur) {
      ++cnt;
    } else {
      if (cur >= 0 && cur < nvalues) atomicAdd(&sizes[cur], (unsigned long long)cnt);
      cur = v;
      cnt = 1;
    }
  }
  if (cur >= 0 && cur < nvalues) atomicAdd(&sizes[cur], (unsigned long long)cnt);
}

}  // namespace mfish

using namespace mfish;

extern "C" int mfish_label_sizes_i32(const int32_t* labels_dev, int64_t n, int32_t nvalues,
                                     unsigned long long* sizes_dev, void* stream) {
  MFISH_REQUIRE(labels_dev && sizes_dev && n >= 0 && nvalues > 0, "label_sizes: bad arguments");
  cudaStream_t st = as_stream(stream);
  MFISH_CUDA(cudaMemsetAsync(sizes_dev, 0, (size_t)nvalues * sizeof(unsigned long long), st));
  if (n == 0) return MFISH_OK;
  const int64_t threads = (n + CCL_RUN - 1) / CCL_RUN;
  label_sizes_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(labels_dev, n, nvalues, sizes_dev);
  count_launch(1);
  return check_launch("label_sizes");
}

extern "C" int mfish_ccl_roots_u8(const uint8_t* mask_dev, int64_t na, int64_t nb, int64_t nc, int connectivity_full,
                                  int32_t* parent_dev, void* stream) {
  MFISH_REQUIRE(mask_dev && parent_dev, "ccl: null pointer");
  MFISH_REQUIRE(na > 0 && nb > 0 && nc > 0, "ccl: empty volume");
  MFISH_REQUIRE(na * nb * nc < (1ll << 31), "ccl: volume too large for int32 labels");
  cudaStream_t st = as_stream(stream);
  const int64_t n = na * nb * nc;
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 32);
  ProfScope prof("ccl", st);
  const int64_t nrows = na * nb;
  ccl_init_rows_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, st>>>(mask_dev, parent_dev, nrows, (int)nc);
  if (connectivity_full)
    ccl_merge_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(parent_dev, (int)na, (int)nb, (int)nc);
  else
    ccl_merge_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(parent_dev, (int)na, (int)nb, (int)nc);
  ccl_flatten_kernel<<<(unsigned)blocks, 256, 0, st>>>(parent_dev, n);
  count_launch(3);
  return check_launch("ccl");
}
