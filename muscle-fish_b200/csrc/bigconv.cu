// Large-radius separable Gaussian passes: the segmentation blurs of the reference,
//   skimage.filters.gaussian(img, (10, 10, 1))  modules/muscle_fish.py:239  (threshold_fiber,  radius 40 in x, y)
//   skimage.filters.gaussian(img, (20, 20, 2))  modules/muscle_fish.py:166  (threshold_nuclei, radius 80 in x, y)
// = scipy.ndimage.gaussian_filter(mode='nearest', truncate=4).  With 81 / 161 taps per output these passes
// are bound by the FP32 FMA pipe (2R+1 FMAs per voxel: 67 G FMA per pass at R = 80 on configs[1], 1.9 ms at
// the 128 FMA/clk/SM of 148 SMs), not by HBM, so the kernels are organised around FMAs per load instead
// of bytes: every thread owns BG_M consecutive outputs of one line and keeps BG_M accumulators AND a
// BG_M-long sliding window of taps in registers, so that one loaded voxel and one loaded tap feed BG_M FMAs
// (the tap of accumulator j at input row r is t[r - p0 - j], i.e. the window of row r-1 shifted by one --
// the loops are unrolled BG_M times and the shift is register renaming).
//   strided axes (x, z): rows come straight from global memory, coalesced over y; the output blocks of a
//                        line are consecutive in launch order, so the (2R+M)/M re-reads hit L2.
//   contiguous axis (y): a 32-row x (256 + 2R) tile is staged in shared memory (odd pitch: lanes = rows).
// Same tap values and boundary rule as the generic kernels; the summation order differs (results agree to
// a few float32 ulps; parity against scipy is the 1e-5 of max|ref| of the other filter tests).
#include <algorithm>

#include "common.cuh"

namespace mfish {

constexpr int BG_M = 16;
constexpr int BG_THREADS = 128;

struct BigArgs {
  const float* in;
  float* out;
  const float* taps;  // device, radius + 1 half taps
  const float* mm;    // normalise on load (may be null)
  int64_t inner, stride, outer_stride, nouter;
  int L, R, bmode;
  int nob;    // output blocks per line = ceil(L / BG_M)
  int64_t nbx;  // column blocks = ceil(inner / BG_THREADS)
};

__device__ __forceinline__ int bg_bmap(int i, int n, int bmode) {
  return bmode == MFISH_BOUNDARY_NEAREST ? bmap<MFISH_BOUNDARY_NEAREST>(i, n) : bmap<MFISH_BOUNDARY_REFLECT>(i, n);
}

__device__ __forceinline__ void bg_norm(const float* mm, float& mn, float& sc, bool& degen) {
  mn = 0.f;
  sc = 1.f;
  degen = false;
  if (mm) {
    const float lo = mm[0], hi = mm[1];
    mn = lo;
    if (hi > lo)
      sc = 1.0f / (hi - lo);
    else
      degen = true;
  }
}

// zero-padded full tap table in shared memory: ext[k + BG_M] = t[|k - R|] for 0 <= k <= 2R, else 0;
// ext has 2R + 1 + 3 * BG_M entries (the sweep runs up to BG_M - 1 rows past the last tap)
__device__ __forceinline__ void bg_fill_taps(float* ext, const float* taps, int R, int nthreads) {
  const int n = 2 * R + 1 + 3 * BG_M;
  for (int k = threadIdx.x; k < n; k += nthreads) {
    const int q = k - BG_M;
    ext[k] = (q >= 0 && q <= 2 * R) ? taps[q >= R ? q - R : R - q] : 0.f;
  }
}

__global__ void __launch_bounds__(BG_THREADS) big_strided_kernel(const BigArgs a) {
  extern __shared__ float bg_ext[];
  bg_fill_taps(bg_ext, a.taps, a.R, BG_THREADS);
  __syncthreads();
  // linear block id: output block fastest, then column block, then outer (plane)
  const int64_t bid = blockIdx.x;
  const int ob = (int)(bid % a.nob);
  const int64_t rest = bid / a.nob;
  const int64_t bx = rest % a.nbx;
  const int64_t outer = rest / a.nbx;
  const int64_t i = bx * BG_THREADS + threadIdx.x;
  if (i >= a.inner) return;
  const float* __restrict__ src = a.in + outer * a.outer_stride + i;
  float* __restrict__ dst = a.out + outer * a.outer_stride + i;
  float mn, sc;
  bool degen;
  bg_norm(a.mm, mn, sc, degen);
  const int p0 = ob * BG_M;
  const int r0 = p0 - a.R;  // first input row of the sweep
  float acc[BG_M], T[BG_M];
#pragma unroll
  for (int j = 0; j < BG_M; ++j) {
    acc[j] = 0.f;
    T[j] = 0.f;
  }
  const int nrows = 2 * a.R + BG_M;  // rows r0 .. p0 + BG_M - 1 + R
#pragma unroll 1
  for (int c0 = 0; c0 < nrows; c0 += BG_M) {
    float v[BG_M];
#pragma unroll
    for (int u = 0; u < BG_M; ++u) {
      const int r = bg_bmap(r0 + c0 + u, a.L, a.bmode);
      v[u] = __ldg(src + (int64_t)r * a.stride);
    }
#pragma unroll
    for (int u = 0; u < BG_M; ++u) {
      const float x = a.mm ? (degen ? 1.0f : (v[u] - mn) * sc) : v[u];
      T[u] = bg_ext[c0 + u + BG_M];  // tap of accumulator 0 at this row; slot u = (c0 + u) % BG_M
#pragma unroll
      for (int j = 0; j < BG_M; ++j) acc[j] = fmaf(T[(u - j + BG_M) % BG_M], x, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < BG_M; ++j)
    if (p0 + j < a.L) dst[(int64_t)(p0 + j) * a.stride] = acc[j];
}

// contiguous axis: block = 32 rows x BGY_CH outputs; thread (lane = row, warp = output block of BG_M)
constexpr int BGY_ROWS = 32;
constexpr int BGY_CH = 256;
constexpr int BGY_THREADS = BGY_ROWS * (BGY_CH / BG_M);

struct BigYArgs {
  const float* in;
  float* out;
  const float* taps;
  const float* mm;
  int64_t nrows;
  int ny, R, bmode, nseg, pitch;
};

__global__ void __launch_bounds__(BGY_THREADS) big_y_kernel(const BigYArgs a) {
  extern __shared__ float bgy_smem[];
  float* ext = bgy_smem;                              // 2R + 1 + 3 BG_M
  float* tile = bgy_smem + (2 * a.R + 1 + 3 * BG_M);  // [BGY_ROWS][pitch]
  bg_fill_taps(ext, a.taps, a.R, BGY_THREADS);
  const int64_t rowblk = blockIdx.x / a.nseg;
  const int seg0 = (int)(blockIdx.x - rowblk * a.nseg) * BGY_CH;
  const int64_t row0 = rowblk * BGY_ROWS;
  float mn, sc;
  bool degen;
  bg_norm(a.mm, mn, sc, degen);
  const int width = BGY_CH + 2 * a.R + BG_M;  // columns seg0 - R .. (the sweep's padded end)
  for (int rr = threadIdx.x >> 5; rr < BGY_ROWS; rr += BGY_THREADS / 32) {
    const int64_t row = row0 + rr;
    if (row >= a.nrows) break;
    const float* __restrict__ src = a.in + row * a.ny;
    for (int c = threadIdx.x & 31; c < width; c += 32) {
      const int y = bg_bmap(seg0 - a.R + c, a.ny, a.bmode);
      const float v = __ldg(src + y);
      tile[rr * a.pitch + c] = a.mm ? (degen ? 1.0f : (v - mn) * sc) : v;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, ob = threadIdx.x >> 5;
  const int64_t row = row0 + lane;
  if (row >= a.nrows) return;
  const int p0 = ob * BG_M;  // first output of this thread inside the segment
  const float* __restrict__ line = tile + lane * a.pitch + p0;  // input column of output p0 at tap 0
  float acc[BG_M], T[BG_M];
#pragma unroll
  for (int j = 0; j < BG_M; ++j) {
    acc[j] = 0.f;
    T[j] = 0.f;
  }
  const int nsteps = 2 * a.R + BG_M;
#pragma unroll 1
  for (int c0 = 0; c0 < nsteps; c0 += BG_M) {
#pragma unroll
    for (int u = 0; u < BG_M; ++u) {
      const float x = line[c0 + u];
      T[u] = ext[c0 + u + BG_M];
#pragma unroll
      for (int j = 0; j < BG_M; ++j) acc[j] = fmaf(T[(u - j + BG_M) % BG_M], x, acc[j]);
    }
  }
  float* __restrict__ dst = a.out + row * a.ny + seg0 + p0;
#pragma unroll
  for (int j = 0; j < BG_M; ++j)
    if (seg0 + p0 + j < a.ny) dst[j] = acc[j];
}

// One op-G pass with a large radius.  taps_dev: radius + 1 half taps on the device.
int big_gauss_pass(const float* in, float* out, int64_t nz, int64_t nx, int64_t ny, int axis, const float* taps_dev,
                   int radius, int boundary, const float* mm, cudaStream_t st) {
  if (axis == MFISH_AXIS_Y) {
    BigYArgs a;
    a.in = in;
    a.out = out;
    a.taps = taps_dev;
    a.mm = mm;
    a.nrows = nz * nx;
    a.ny = (int)ny;
    a.R = radius;
    a.bmode = boundary;
    a.nseg = (int)((ny + BGY_CH - 1) / BGY_CH);
    a.pitch = (BGY_CH + 2 * radius + BG_M) | 1;  // odd: the 32 rows of a warp fall into 32 banks
    const size_t smem = (size_t)(2 * radius + 1 + 3 * BG_M + BGY_ROWS * a.pitch) * sizeof(float);
    MFISH_REQUIRE(smem <= 200 * 1024, "gaussian: radius too large for the tiled y pass");
    MFISH_CUDA(cudaFuncSetAttribute(big_y_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t blocks = ((a.nrows + BGY_ROWS - 1) / BGY_ROWS) * a.nseg;
    MFISH_REQUIRE(blocks < (1ll << 31), "gaussian(y): grid too large");
    ProfScope prof("conv_big_y", st);
    big_y_kernel<<<(unsigned)blocks, BGY_THREADS, smem, st>>>(a);
    count_launch(1);
    return check_launch("conv_big_y");
  }
  BigArgs a;
  a.in = in;
  a.out = out;
  a.taps = taps_dev;
  a.mm = mm;
  if (axis == MFISH_AXIS_Z) {
    a.inner = nx * ny;
    a.stride = nx * ny;
    a.outer_stride = 0;
    a.nouter = 1;
    a.L = (int)nz;
  } else {
    a.inner = ny;
    a.stride = ny;
    a.outer_stride = nx * ny;
    a.nouter = nz;
    a.L = (int)nx;
  }
  a.R = radius;
  a.bmode = boundary;
  a.nob = (a.L + BG_M - 1) / BG_M;
  a.nbx = (a.inner + BG_THREADS - 1) / BG_THREADS;
  const int64_t blocks = a.nouter * a.nbx * a.nob;
  MFISH_REQUIRE(blocks < (1ll << 31), "gaussian(strided): grid too large");
  const size_t smem = (size_t)(2 * radius + 1 + 3 * BG_M) * sizeof(float);
  ProfScope prof("conv_big_strided", st);
  big_strided_kernel<<<(unsigned)blocks, BG_THREADS, smem, st>>>(a);
  count_launch(1);
  return check_launch("conv_big_strided");
}

}  // namespace mfish
